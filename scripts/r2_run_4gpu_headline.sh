#!/bin/bash
mkdir -p gpurun_out
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29515 bench.py --gpus 4 --steps 10 --warmup 3 --no-extra > gpurun_out/r02_bench_4gpu_headline.json 2>/dev/null; echo rc=$?
python -c "
import json; d=json.load(open('gpurun_out/r02_bench_4gpu_headline.json')); print('4gpu round_ms', d['ms_per_step'], 'value', d['value'], 'e2e', d['e2e']['value'])"
