#!/bin/bash
# One-GPU round: GPU suite, bench + reference arm, phase breakdowns, the RM_CHECKED pass.   tools/gpu_round.sh TAG
TAG=${1:-x}
timeout -s KILL 900 python -m pytest tests -m gpu -x -q -rs > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_$TAG.log
timeout -s KILL 400 python bench.py --steps 200 --warmup 10 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"; tail -2 gpurun_out/bench_$TAG.err
timeout -s KILL 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_${TAG}_reference.json 2> gpurun_out/bench_${TAG}_reference.err; echo "bench reference rc=$?"
for w in cornell_4k demo dodecahedron_4k cornell_1080p; do timeout -s KILL 120 python tools/run_phases.py $w 8 2>&1 | grep "frame [7]" | sed "s/^/$w: /"; done | tee gpurun_out/phases_$TAG.txt
tools/gpu_checked.sh $TAG
timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$TAG.txt 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke_$TAG.txt
