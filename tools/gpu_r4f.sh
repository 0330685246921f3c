#!/bin/bash
TAG=${1:-r4f}
RM_B200_DELIVERY_TRACE=1 timeout -s KILL 300 python tools/e2e_breakdown.py 2> gpurun_out/e2e_trace_full_$TAG.txt | tee gpurun_out/e2e_breakdown_$TAG.txt
awk 'NR>=130 && NR<=136' gpurun_out/e2e_trace_full_$TAG.txt
grep -n "retained" gpurun_out/e2e_trace_full_$TAG.txt | awk -F: '$1>=340 && $1<=346'
