#!/bin/bash
# One GPU iteration: parity tests, the phase breakdown of every workload, the bench (both arms), then -- only after those
# exited 0 without ncu -- the ncu launch list of the bench command and one ncu --set full capture of the render kernel.
# Every step runs under a hard time-out (timeout -s KILL): a kernel that never retires blocks the host in
# cudaStreamSynchronize and must cost seconds of the GPU budget, not the whole call (lesson of cycle r3a: 7 minutes).
# Usage: tools/gpu_cycle.sh TAG [noncu]
TAG=${1:-x}
T="timeout -s KILL"
$T 180 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_$TAG.log
for w in cornell_4k demo cornell_1080p dodecahedron_4k stress_4k_bvh; do $T 60 python tools/run_phases.py $w 8 2>&1 | grep "frame 7" | sed "s/^/$w: /"; done | tee gpurun_out/phases_$TAG.log
$T 240 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_$TAG.log 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"
$T 240 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref_$TAG.log 2> gpurun_out/bench_ref_$TAG.err; echo "bench reference rc=$?"
for wl in stress_4k_bvh stress_8k_bvh; do
  $T 90 python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${TAG}_$wl.log 2> gpurun_out/bench_${TAG}_$wl.err; echo "bench $wl rc=$?"
done
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/bench_$TAG*.log")):
    l = [x for x in open(f) if x.startswith("{")]
    if l:
        d = json.loads(l[-1]); r = d["roofline"]
        print(f, "step ms %.4f kernels ms %.4f frac %s issue %s e2e ms %.3f cpu ms %s" % (d["ms_per_step"], r["kernel_ms"], r["frac"], r["issue_slot_frac"], d["e2e"]["ms_per_frame"], d.get("cpu_baseline") and d["cpu_baseline"].get("ms_per_frame")))
PY
if [ "$2" != "noncu" ]; then
$T 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches_$TAG.log 2>&1; echo "ncu launch list rc=$?"
$T 300 ncu --set full --clock-control none --import-source on --kernel-name regex:render_fast -c 2 -f -o gpurun_out/prof_$TAG python tools/run_phases.py cornell_4k 3 > gpurun_out/ncu_$TAG.log 2>&1; echo "ncu full rc=$?"
fi
