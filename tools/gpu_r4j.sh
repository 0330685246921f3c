#!/bin/bash
TAG=${1:-r4j}
RM_B200_DELIVERY_TRACE=1 timeout -s KILL 300 python tools/e2e_breakdown.py 2> gpurun_out/e2e_trace_full_$TAG.txt | tee gpurun_out/e2e_breakdown_$TAG.txt
grep -A4 "== retained resident" gpurun_out/e2e_trace_full_$TAG.txt | cut -c1-220 | tail -3
grep -A40 "== retained with upload" gpurun_out/e2e_trace_full_$TAG.txt | cut -c1-220 | tail -3
