cd any-stereo_b200/csrc
for v in new base new base; do
  cp lib_$v.so.keep libanystereo_b200.so
  (cd ../..; echo "== $v"; python tools/train_step.py --steps 8 --engine bf16x3 2>/dev/null | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print([round(x,1) for x in d['ms_per_step'][-4:]], d.get('phase_ms_median')['backward'])"; python tools/microbench.py --only train 2>&1 | grep "tcgen05" | cut -c1-80)
done
cp lib_new.so.keep libanystereo_b200.so
