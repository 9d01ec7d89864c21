# Evidence run of round 2 (GPU box): bench line, microbench, ncu launch list + --set full captures, drop-in EPE
set -x
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r02_final2.json 2> gpurun_out/bench_r02_final2.err
python tools/microbench.py --json gpurun_out/microbench_r02_final2.json > gpurun_out/mb2.log 2>&1
AS_HOTLOOP_GRAPH=0 timeout 600 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_r02_final2.csv python bench.py --eager --steps 1 --warmup 3 --no-cpu-baseline --no-other-configs > gpurun_out/ncu_l.log 2>&1
AS_HOTLOOP_GRAPH=0 timeout 600 ncu --set full --import-source on --clock-control none -k regex:geo_lookup_convc1_tap -s 40 -c 1 -o gpurun_out/prof_c1tap_final -f python bench.py --eager --steps 1 --warmup 3 --no-cpu-baseline --no-other-configs > gpurun_out/ncu_t.log 2>&1
AS_HOTLOOP_GRAPH=0 timeout 600 ncu --set full --clock-control none -k regex:conv_umma_kernel -s 300 -c 10 -o gpurun_out/prof_conv_final -f python bench.py --eager --steps 1 --warmup 3 --no-cpu-baseline --no-other-configs > gpurun_out/ncu_c.log 2>&1
python tools/dropin.py --engines default,bf16x3,fp16 --json gpurun_out/dropin_r02_final2.json > gpurun_out/dropin_r02_final2.log 2>&1
ls -la gpurun_out | tail -12
