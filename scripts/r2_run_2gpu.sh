#!/bin/bash
# 2 GPUs: the sharded-round parity tests (both collectives) and the peer-memory FedAvg test, then the headline under torchrun
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 900 python -m pytest tests/test_gpu_round2.py -q -m gpu -k "two_gpu or peer_fedavg" 2>&1 | tail -6 > gpurun_out/r02_2gpu_tests.log
tail -4 gpurun_out/r02_2gpu_tests.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-extra > gpurun_out/r02_bench_2gpu_headline.json 2> gpurun_out/r02_bench_2gpu.err
echo "bench rc=$?"
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02_bench_2gpu_headline.json"))
print("2gpu round_ms", d["ms_per_step"], "value", d["value"], "e2e", d["e2e"]["value"])
PY
