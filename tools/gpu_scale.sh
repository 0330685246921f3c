#!/bin/bash
# Scaling run (gpurun --gpus N): cross-GPU parity test at world N, bench at 1, 2, 4, 8 GPUs on the headline workload and on
# the compute-heavy stress workload, and the frame kernel's phase breakdown at N.   tools/gpu_scale.sh TAG N
TAG=${1:-x}; N=${2:-8}
nvidia-smi topo -m > gpurun_out/topo_$TAG.log 2>&1
timeout -s KILL 400 python -m pytest tests/test_gpu_parity.py -x -q -k "peer_exchange_across_gpus" > gpurun_out/pytest_$TAG.log 2>&1; echo "cross-GPU pytest rc=$?"; tail -2 gpurun_out/pytest_$TAG.log
run() {  # workload n extra-args
  if [ $2 -eq 1 ]; then
    timeout -s KILL 400 python bench.py --workload $1 --no-cpu-baseline $3 > gpurun_out/bench_${TAG}_$1_n$2.log 2> gpurun_out/bench_${TAG}_$1_n$2.err
  else
    timeout -s KILL 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $2 --master-addr 127.0.0.1 --master-port 2951$2 bench.py --workload $1 --gpus $2 $3 > gpurun_out/bench_${TAG}_$1_n$2.log 2> gpurun_out/bench_${TAG}_$1_n$2.err
  fi
  echo "bench $1 n$2 rc=$?"
}
for n in 1 2 4 8; do
  if [ $n -le $N ]; then
    run cornell_4k $n ""
    run stress_4k $n "--steps 5 --warmup 3"
  fi
done
timeout -s KILL 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29520 tools/run_phases.py cornell_4k 8 > gpurun_out/phases_${TAG}_n$N.log 2>&1
grep "frame 7" gpurun_out/phases_${TAG}_n$N.log
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/bench_${TAG}_*_n*.log")):
    l=[x for x in open(f) if x.startswith("{")]
    if l:
        d=json.loads(l[-1]); r=d["roofline"]
        print(f, "n=%d step ms %.4f kernels %.4f frac %.3f issue %.3f e2e ms %.3f exch %s" % (d["n_gpus"], d["ms_per_step"], r["kernel_ms"], r["frac"], r["issue_slot_frac"], d["e2e"]["ms_per_frame"], d["config"].get("exchange")))
PY
