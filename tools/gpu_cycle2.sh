#!/bin/bash
# Multi-GPU iteration (gpurun --gpus N): parity tests incl. the cross-GPU peer exchange, then bench at 1..N GPUs.
# Usage: tools/gpu_cycle2.sh TAG N
TAG=${1:-x}; N=${2:-2}
nvidia-smi topo -m > gpurun_out/topo_$TAG.log 2>&1
timeout -s KILL 300 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_$TAG.log
timeout -s KILL 300 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_${TAG}_n1.log 2> gpurun_out/bench_${TAG}_n1.err; echo "bench n1 rc=$?"
for n in 2 4 8; do
  if [ $n -le $N ]; then
    timeout -s KILL 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/bench_${TAG}_n$n.log 2> gpurun_out/bench_${TAG}_n$n.err; echo "bench n$n rc=$?"
    tail -3 gpurun_out/bench_${TAG}_n$n.err
  fi
done
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/bench_${TAG}_n*.log")):
    l=[x for x in open(f) if x.startswith("{")]
    if l:
        d=json.loads(l[-1]); r=d["roofline"]
        print(f, "n=%d step ms %.4f k0 %.4f k1 %.4f frac %.3f issue %.3f e2e ms %.3f exch %s" % (d["n_gpus"], d["ms_per_step"], r["prepare_kernel_ms"], r["kernel_ms"], r["frac"], r["issue_slot_frac"], d["e2e"]["ms_per_frame"], d["config"].get("exchange")))
PY
