#!/bin/bash
# A/B of build variants on one GPU: tools/gpu_ab.sh TAG "<flags A>" "<flags B>" ...   (flags: RM_NVCC_EXTRA values, "-" = none)
TAG=$1; shift
for v in "$@"; do
  if [ "$v" = "-" ]; then export RM_NVCC_EXTRA=""; else export RM_NVCC_EXTRA="$v"; fi
  python -m rusty_marcher_b200.build --force > /dev/null 2>&1 || echo "build failed for $v"
  echo "== variant: $v"
  timeout -s KILL 300 python -m pytest tests/test_gpu_parity.py -x -q -k "north_star or frame_level_call_equals" 2>&1 | tail -1
  timeout -s KILL 120 python tools/run_phases.py cornell_4k 10 2>&1 | grep "frame [6-9]"
  timeout -s KILL 120 python tools/run_phases.py demo 8 2>&1 | grep "frame [7]"
  timeout -s KILL 120 python tools/run_phases.py dodecahedron_4k 8 2>&1 | grep "frame [7]"
done 2>&1 | tee gpurun_out/ab_$TAG.log
export RM_NVCC_EXTRA=""
