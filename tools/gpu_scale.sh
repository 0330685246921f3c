#!/bin/bash
# Scaling run (gpurun --gpus N): bench at 1, 2, 4, 8 GPUs + the frame kernel's phase breakdown at N.   tools/gpu_scale.sh TAG N
TAG=${1:-x}; N=${2:-8}
nvidia-smi topo -m > gpurun_out/topo_$TAG.log 2>&1
for n in 1 2 4 8; do
  if [ $n -le $N ]; then
    if [ $n -eq 1 ]; then
      timeout 300 python bench.py  --no-cpu-baseline > gpurun_out/bench_${TAG}_n$n.log 2> gpurun_out/bench_${TAG}_n$n.err
    else
      timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n  > gpurun_out/bench_${TAG}_n$n.log 2> gpurun_out/bench_${TAG}_n$n.err
    fi
    echo "bench n$n rc=$?"
  fi
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29520 tools/run_phases.py cornell_4k 8 > gpurun_out/phases_${TAG}_n$N.log 2>&1
grep "frame 7" gpurun_out/phases_${TAG}_n$N.log
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/bench_${TAG}_n*.log")):
    l=[x for x in open(f) if x.startswith("{")]
    if l:
        d=json.loads(l[-1]); r=d["roofline"]
        print(f, "n=%d step ms %.4f kernels %.4f frac %.3f issue %.3f e2e ms %.3f exch %s" % (d["n_gpus"], d["ms_per_step"], r["kernel_ms"], r["frac"], r["issue_slot_frac"], d["e2e"]["ms_per_frame"], d["config"].get("exchange")))
PY
