#!/bin/bash
# One GPU iteration for the hierarchy path (RmParams.accel): parity tests, the stress workloads with and without it,
# then the headline bench as a regression check.
# Usage: tools/gpu_bvh.sh TAG
TAG=${1:-x}
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_$TAG.log
for wl in stress_4k_bvh stress_8k_bvh; do
  timeout 300 python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${TAG}_$wl.log 2> gpurun_out/bench_${TAG}_$wl.err; echo "bench $wl rc=$?"; tail -2 gpurun_out/bench_${TAG}_$wl.err
done
timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/bench_${TAG}.log 2> gpurun_out/bench_${TAG}.err; echo "bench rc=$?"
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/bench_${TAG}*.log")):
    l=[x for x in open(f) if x.startswith("{")]
    if l:
        d=json.loads(l[-1]); r=d["roofline"]
        print(f, "step ms %.4f kernels ms %.4f segs %d value %.3e frac %s brute ms %s e2e ms %.3f clocks %s" % (d["ms_per_step"], r["kernel_ms"], d["config"]["segments_per_frame"], d["value"], r["frac"], r.get("brute_force_ms_same_frame"), d["e2e"]["ms_per_frame"], d["clocks"]))
PY
