#!/bin/bash
TAG=${1:-r4d}
timeout -s KILL 600 python -m pytest tests -m gpu -x -q -k "host_delivery or frame_level or full_size" > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_$TAG.log
timeout -s KILL 300 python tools/e2e_breakdown.py 2>&1 | tee gpurun_out/e2e_breakdown_$TAG.txt
RM_B200_DELIVERY_TRACE=1 timeout -s KILL 300 python tools/e2e_breakdown.py 2>&1 | grep "rm delivery" | awk 'NR%40==1' | head -12 | tee gpurun_out/e2e_trace_$TAG.txt
