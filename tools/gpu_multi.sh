#!/bin/bash
# multi-GPU call (gpurun --gpus N): the multi-GPU tests, bench at N (torchrun), phases at N
N=${1:-2}; TAG=${2:-r4n2}
nvidia-smi -L | head -8
timeout -s KILL 600 python -m pytest tests -m gpu -x -q -rs -k "peer_exchange" > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_$TAG.log
timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 200 --warmup 10 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench N=$N rc=$?"; tail -5 gpurun_out/bench_$TAG.err
timeout -s KILL 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 tools/run_phases.py cornell_4k 8 2>&1 | grep "frame [67]" | tee gpurun_out/phases_$TAG.txt
