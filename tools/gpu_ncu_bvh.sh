#!/bin/bash
# ncu --set full capture of the hierarchy kernel on the stress_4k_bvh frame (the program ran without ncu in cycle r3e)
TAG=${1:-x}
timeout -s KILL 55 ncu --set full --clock-control none --import-source on --kernel-name regex:render_fast -s 2 -c 1 -f -o gpurun_out/prof_$TAG python tools/run_phases.py stress_4k_bvh 3 > gpurun_out/ncu_$TAG.log 2>&1; echo "ncu full rc=$?"
tail -3 gpurun_out/ncu_$TAG.log
