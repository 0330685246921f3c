#!/bin/bash
# Usage: tools/gpu_diag.sh TAG -- checks the hierarchy path against brute force (tools/diag_bvh2.py: primary ids and floats,
# twice, at two sizes; each in its own process under a hard time-out); when it agrees: GPU test suite + benches
TAG=${1:-x}
LOG=gpurun_out/diag_$TAG.log
: > $LOG
ok=1
timeout -s KILL 12 python tools/diag_bvh2.py 320 192 1 _$TAG >> $LOG 2>&1 || ok=0
timeout -s KILL 15 python tools/diag_bvh2.py 1280 704 6 _$TAG >> $LOG 2>&1 || ok=0
echo "hierarchy == brute force: ok=$ok" >> $LOG
if [ $ok -eq 1 ]; then
  timeout -s KILL 70 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?" >> $LOG; tail -3 gpurun_out/pytest_$TAG.log >> $LOG
  timeout -s KILL 30 python bench.py --workload stress_4k_bvh --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${TAG}_stress_4k_bvh.log 2> gpurun_out/bench_${TAG}_stress_4k_bvh.err; echo "bench stress_4k_bvh rc=$?" >> $LOG
  timeout -s KILL 45 python bench.py --workload stress_8k_bvh --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${TAG}_stress_8k_bvh.log 2> gpurun_out/bench_${TAG}_stress_8k_bvh.err; echo "bench stress_8k_bvh rc=$?" >> $LOG
  timeout -s KILL 30 python bench.py --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/bench_${TAG}.log 2> gpurun_out/bench_${TAG}.err; echo "bench rc=$?" >> $LOG
fi
cat $LOG
