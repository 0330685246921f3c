#!/bin/bash
TAG=${1:-r4c}
timeout -s KILL 900 python -m pytest tests -m gpu -x -q -rs > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_$TAG.log
for w in cornell_4k demo stress_4k_bvh stress_8k_bvh; do timeout -s KILL 120 python tools/run_phases.py $w 8 2>&1 | grep "frame [7]"; done | tee gpurun_out/phases_$TAG.txt
timeout -s KILL 300 python tools/e2e_breakdown.py 2>&1 | tee gpurun_out/e2e_breakdown_$TAG.txt
RM_B200_DELIVERY_TRACE=1 timeout -s KILL 300 python tools/e2e_breakdown.py 2>&1 | grep "rm delivery" | awk 'NR%40==1' | head -12 | tee gpurun_out/e2e_trace_$TAG.txt
timeout -s KILL 400 python bench.py --steps 200 --warmup 10 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_$TAG.err
