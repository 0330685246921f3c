#!/bin/bash
# ncu evidence of one workload: launch list of the bench command + one --set full capture of the render kernel.
# Usage: tools/gpu_ncu.sh TAG WORKLOAD   (one GPU; run only after the same bench command has exited 0 without ncu)
TAG=${1:-x}; WL=${2:-cornell_4k}
timeout -s KILL 300 python bench.py --workload $WL --steps 3 --warmup 3 --no-cpu-baseline --heavy '' > gpurun_out/ncu_plain_$TAG.json 2> gpurun_out/ncu_plain_$TAG.err || { echo "plain run failed"; exit 1; }
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --workload $WL --steps 3 --warmup 3 --no-cpu-baseline --heavy '' > gpurun_out/ncu_list_$TAG.log 2>&1; echo "launch list rc=$?"
timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k regex:render_fast_kernel -s 4 -c 2 -o gpurun_out/prof_$TAG -f python bench.py --workload $WL --steps 3 --warmup 3 --no-cpu-baseline --heavy '' > gpurun_out/ncu_full_$TAG.log 2>&1; echo "full capture rc=$?"
ls -la gpurun_out/prof_$TAG.ncu-rep
