# Experiment (run on the GPU box): rebuild the tap-major fused lookup kernel with ablation flags and time it.
#   AS_TAP_NOLOAD: no gather traffic (addresses still computed);  AS_TAP_NOEPI: no output stores
# r02 result at config 2 (bf16x3 planes): full 93 us, no loads 50, no stores 77, neither 42 -- i.e. loads and compute
# do not overlap well (LDG issue blocks in lg_throttle), and 16-byte-per-lane output stores cost 16-20 us (32-byte
# stores: 10).
for F in "" "-DAS_TAP_NOLOAD" "-DAS_TAP_NOEPI" "-DAS_TAP_NOLOAD -DAS_TAP_NOEPI"; do
  touch any-stereo_b200/csrc/lookup_c1_tap.cu; make -C any-stereo_b200/csrc -s -j8 EXTRA="$F" 2>&1 | tail -2
  echo "== $F"; timeout 300 python tools/microbench.py --only lookup 2>&1 | grep -i "fused.*_tap" | cut -c1-90
done
touch any-stereo_b200/csrc/lookup_c1_tap.cu; make -C any-stereo_b200/csrc -s -j8
