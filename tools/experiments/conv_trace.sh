# Experiment (GPU box): where does the MMA-issuing thread of the convolution kernel wait?  Builds the library with
# -DAS_CONV_TRACE, runs ONE eager iteration of the config-2 update block and prints one line per convolution launch.
touch any-stereo_b200/csrc/conv_umma.cu; make -C any-stereo_b200/csrc -s -j8 EXTRA=-DAS_CONV_TRACE 2>&1 | tail -2
AS_CONV_TRACE_PRINT=1 AS_HOTLOOP_GRAPH=0 python bench.py --eager --steps 1 --warmup 3 --no-cpu-baseline --no-other-configs 2>&1 | grep "^conv N" | tail -14
touch any-stereo_b200/csrc/conv_umma.cu; make -C any-stereo_b200/csrc -s -j8 2>&1 | tail -2
