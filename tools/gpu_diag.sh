#!/bin/bash
# Usage: tools/gpu_diag.sh TAG -- checks the hierarchy path against brute force; when it agrees: GPU test suite + benches
TAG=${1:-x}
LOG=gpurun_out/diag_$TAG.log
: > $LOG
ok=1
timeout -s KILL 12 python tools/diag_bvh2.py 320 192 1 _v1 >> $LOG 2>&1 || ok=0
timeout -s KILL 15 python tools/diag_bvh2.py 1280 704 6 _v1 >> $LOG 2>&1 || ok=0
echo "v1 ok=$ok" >> $LOG
if [ $ok -eq 1 ]; then
  timeout -s KILL 70 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?" >> $LOG; tail -3 gpurun_out/pytest_$TAG.log >> $LOG
  timeout -s KILL 30 python bench.py --workload stress_4k_bvh --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${TAG}_stress_4k_bvh.log 2> gpurun_out/bench_${TAG}_stress_4k_bvh.err; echo "bench stress_4k_bvh rc=$?" >> $LOG
  timeout -s KILL 45 python bench.py --workload stress_8k_bvh --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${TAG}_stress_8k_bvh.log 2> gpurun_out/bench_${TAG}_stress_8k_bvh.err; echo "bench stress_8k_bvh rc=$?" >> $LOG
  timeout -s KILL 30 python bench.py --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/bench_${TAG}.log 2> gpurun_out/bench_${TAG}.err; echo "bench rc=$?" >> $LOG
else
  cp tools/variants/librm_b200_v2.so rusty_marcher_b200/librm_b200.so
  timeout -s KILL 12 python tools/diag_bvh2.py 320 192 1 _v2 >> $LOG 2>&1; echo "v2 small rc=$?" >> $LOG
  timeout -s KILL 15 python tools/diag_bvh2.py 1280 704 6 _v2 >> $LOG 2>&1; echo "v2 large rc=$?" >> $LOG
fi
cat $LOG
