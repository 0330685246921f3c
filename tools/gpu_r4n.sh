#!/bin/bash
TAG=${1:-r4n}
timeout -s KILL 120 python -m pytest tests -m gpu -x -q -k "fp32_kernels_meet or hierarchy_changes_no_pixel" > gpurun_out/pytest_${TAG}_quick.log 2>&1; echo "quick pytest rc=$?"; tail -3 gpurun_out/pytest_${TAG}_quick.log
timeout -s KILL 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_$TAG.log
for w in demo stress_4k_bvh; do timeout -s KILL 120 python tools/run_phases.py $w 8 2>&1 | grep "frame [7]"; done | tee gpurun_out/phases_$TAG.txt
BANDS_ALL=1 BANDS_STRIDES=1,8 timeout -s KILL 200 python tools/run_bands.py stress_8k_bvh 5 2>&1 | grep "stride" | head -4 | tee gpurun_out/bands_$TAG.txt
BANDS_ALL=1 BANDS_STRIDES=1 timeout -s KILL 200 python tools/run_bands.py stress_4k 4 2>&1 | grep "stride" | tee -a gpurun_out/bands_$TAG.txt
