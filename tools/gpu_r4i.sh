#!/bin/bash
TAG=${1:-r4i}
timeout -s KILL 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_$TAG.log
timeout -s KILL 400 python bench.py --steps 200 --warmup 10 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"; tail -2 gpurun_out/bench_$TAG.err
for wl in stress_4k_bvh stress_8k_bvh; do
timeout -s KILL 400 python bench.py --workload $wl --steps 10 --warmup 3 --heavy '' --no-cpu-baseline > gpurun_out/bench_${TAG}_$wl.json 2> gpurun_out/bench_${TAG}_$wl.err; echo "bench $wl rc=$?"; tail -2 gpurun_out/bench_${TAG}_$wl.err
done
