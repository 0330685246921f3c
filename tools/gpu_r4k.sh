#!/bin/bash
TAG=${1:-r4k}
for t in 2 4 8 12 16; do echo "== RM_B200_HOST_THREADS=$t"; RM_B200_HOST_THREADS=$t timeout -s KILL 300 python tools/e2e_breakdown.py 2>/dev/null | grep -E "resident|upload \+ render|f64"; done | tee gpurun_out/e2e_threads_$TAG.txt
