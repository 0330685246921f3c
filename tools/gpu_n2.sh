#!/bin/bash
TAG=${1:-x}
timeout -s KILL 300 python tools/frame_rate.py cornell_4k 3000 2>&1 | tee gpurun_out/frame_rate_$TAG.txt
tools/gpu_multi.sh 2 ${TAG}_n2
