#!/bin/bash
# Quick check of a kernel change on one GPU: suite, phase times of the four scene classes, one GPU's share of the heavy frame
TAG=${1:-x}
timeout -s KILL 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_$TAG.log
for w in cornell_4k demo stress_4k_bvh stress_8k_bvh; do timeout -s KILL 120 python tools/run_phases.py $w 8 2>&1 | grep "frame [7]" | sed "s/^/$w: /"; done | tee gpurun_out/phases_$TAG.txt
BANDS_ALL=1 BANDS_STRIDES=8 timeout -s KILL 200 python tools/run_bands.py stress_8k_bvh 5 2>&1 | grep "stride" | head -3 | tee gpurun_out/bands_$TAG.txt
BANDS_ALL=1 BANDS_STRIDES=1 timeout -s KILL 200 python tools/run_bands.py stress_4k 4 2>&1 | grep "stride" | tee -a gpurun_out/bands_$TAG.txt
