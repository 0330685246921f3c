#!/bin/bash
# round-2 first GPU call: GPU test suite, bench, phases, sanitizer, A/B of the opaque instantiation's occupancy target
TAG=${1:-r4a}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader > gpurun_out/${TAG}_gpu.txt 2>&1
timeout -s KILL 900 python -m pytest tests -m gpu -x -q -rs > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_$TAG.log
timeout -s KILL 200 python bench.py --steps 200 --warmup 10 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"
for w in cornell_4k demo dodecahedron_4k; do timeout -s KILL 120 python tools/run_phases.py $w 8 2>&1 | grep "frame [7]"; done | tee gpurun_out/phases_$TAG.txt
tools/sanitize.sh $TAG
tools/gpu_ab.sh $TAG "-DRM_K1_MIN_BLOCKS_OPAQUE=2" "-DRM_K1_MIN_BLOCKS_OPAQUE=3" "-DRM_K1_MIN_BLOCKS_OPAQUE=4"
