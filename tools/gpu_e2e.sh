#!/bin/bash
# Host-delivery changes on one GPU: suite, bench (e2e figures), the e2e breakdown with the library's delivery trace
TAG=${1:-x}
timeout -s KILL 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_$TAG.log
timeout -s KILL 400 python bench.py --steps 200 --warmup 10 --no-cpu-baseline > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"; tail -2 gpurun_out/bench_$TAG.err
python -c "
import json; d=json.loads(open('gpurun_out/bench_$TAG.json').read().strip().splitlines()[-1]); print('ms_per_step', d['ms_per_step'], 'e2e', {k: v for k, v in d['e2e'].items() if k.endswith('ms_per_frame') or k == 'frame_matches_n1'})"
RM_B200_DELIVERY_TRACE=1 timeout -s KILL 300 python tools/e2e_breakdown.py > gpurun_out/e2e_breakdown_$TAG.txt 2> gpurun_out/e2e_trace_$TAG.txt; echo "breakdown rc=$?"; cat gpurun_out/e2e_breakdown_$TAG.txt; grep "rm delivery" gpurun_out/e2e_trace_$TAG.txt | sed -n '20,24p;200,204p'
