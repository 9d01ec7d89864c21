# Experiment (GPU box): same-box A/B of convolution-kernel knobs on the headline step
for env in "AS_CONV_WIDE=1" "AS_CONV_WIDE=0" "AS_CONV_WIDE=1" "AS_CONV_WIDE=0"; do
echo "== $env"
env $env python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-other-configs 2>/dev/null | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print({k:d[k] for k in ['value','ms_per_step']}, d['e2e']['value'], d['clocks']['sm_mhz'], d['roofline_update_block']['avg_us_per_iteration'])"
done
