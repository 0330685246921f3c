#!/bin/bash
# Host-side changes on one GPU box: suite, bench (headline + graph A/B), first-frame breakdown of the 104k-primitive scene
TAG=${1:-x}
timeout -s KILL 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_$TAG.log
timeout -s KILL 400 python bench.py --steps 200 --warmup 10 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"; tail -2 gpurun_out/bench_$TAG.err
timeout -s KILL 300 python tools/ingest_breakdown.py > gpurun_out/ingest_$TAG.txt 2>&1; echo "ingest rc=$?"; cat gpurun_out/ingest_$TAG.txt
timeout -s KILL 300 python bench.py --workload stress_8k_bvh --steps 5 --warmup 3 --heavy '' --no-cpu-baseline > gpurun_out/bench_${TAG}_stress8k.json 2> gpurun_out/bench_${TAG}_stress8k.err; echo "bench stress rc=$?"
