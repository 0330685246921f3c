#!/bin/bash
TAG=${1:-r4g}
timeout -s KILL 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_$TAG.log
for g in 1 0 1 0; do echo "== RM_B200_GRAPH=$g"; RM_B200_GRAPH=$g timeout -s KILL 120 python tools/run_phases.py cornell_4k 12 2>&1 | grep "frame \(8\|9\|10\|11\)"; done | tee gpurun_out/graph_ab_$TAG.txt
for g in 1 0; do echo "== bench RM_B200_GRAPH=$g"; RM_B200_GRAPH=$g timeout -s KILL 200 python bench.py --steps 300 --warmup 10 --no-cpu-baseline --heavy '' 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['frame_matches_n1'])"; done | tee -a gpurun_out/graph_ab_$TAG.txt
