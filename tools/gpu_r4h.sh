#!/bin/bash
TAG=${1:-r4h}
for v in "-" "-DRM_GLASS64_INLINE"; do
  if [ "$v" = "-" ]; then export RM_NVCC_EXTRA=""; else export RM_NVCC_EXTRA="$v"; fi
  python -m rusty_marcher_b200.build --force --verbose 2>&1 | grep -E "Li2EEE" -A2 | grep -E "spill|registers" | head -6
  echo "== variant: $v"
  BANDS_ALL=1 BANDS_STRIDES=1,8 timeout -s KILL 200 python tools/run_bands.py stress_8k_bvh 5 2>&1 | grep "stride"
  BANDS_ALL=1 BANDS_STRIDES=1 timeout -s KILL 200 python tools/run_bands.py stress_4k_bvh 6 2>&1 | grep "stride"
done 2>&1 | tee gpurun_out/glass_inline_ab_$TAG.txt
export RM_NVCC_EXTRA=""
