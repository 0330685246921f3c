#!/bin/bash
# round-2 second GPU call: suite, bench with frame check + heavy leg + e2e variants, phases, cooperative-launch A/B
TAG=${1:-r4b}
timeout -s KILL 900 python -m pytest tests -m gpu -x -q -rs > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_$TAG.log
timeout -s KILL 400 python bench.py --steps 200 --warmup 10 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_$TAG.err
for w in cornell_4k demo dodecahedron_4k stress_4k_bvh; do timeout -s KILL 120 python tools/run_phases.py $w 8 2>&1 | grep "frame [7]"; done | tee gpurun_out/phases_$TAG.txt
echo "== cooperative off"; RM_B200_COOPERATIVE=0 timeout -s KILL 120 python tools/run_phases.py cornell_4k 10 2>&1 | grep "frame [6-9]" | tee -a gpurun_out/phases_$TAG.txt
echo "== host threads 4 / 8"; for t in 4 8; do RM_B200_HOST_THREADS=$t timeout -s KILL 200 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --heavy '' 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['e2e'])"; done | tee gpurun_out/e2e_threads_$TAG.txt
nproc >> gpurun_out/e2e_threads_$TAG.txt; lscpu | grep -E "Model name|Socket|Thread|Core" >> gpurun_out/e2e_threads_$TAG.txt
