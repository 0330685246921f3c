#!/bin/bash
TAG=${1:-r4l}
timeout -s KILL 300 python tools/e2e_breakdown.py 2>/dev/null | tee gpurun_out/e2e_breakdown_$TAG.txt
timeout -s KILL 600 python -m pytest tests -m gpu -x -q -k "host_delivery or walk_stat" 2>&1 | tail -2
timeout -s KILL 400 python bench.py --steps 200 --warmup 10 --heavy '' > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"
python -c "
import json; d=json.loads(open('gpurun_out/bench_$TAG.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['e2e'])"
