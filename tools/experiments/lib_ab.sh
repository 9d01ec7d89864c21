# Experiment (GPU box): same-box A/B of two prebuilt libraries (csrc/lib_base.so.keep vs csrc/lib_new.so.keep) on the headline step
cd any-stereo_b200/csrc
for v in new base new base; do
  cp lib_$v.so.keep libanystereo_b200.so
  echo "== $v"
  (cd ../..; python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-other-configs 2>/dev/null | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print({k:d[k] for k in ['value','ms_per_step']}, d['e2e']['value'], d['clocks']['sm_mhz'])")
done
cp lib_new.so.keep libanystereo_b200.so
