#!/bin/bash
# 2+ GPUs: peer-memory FedAvg tests and the collective micro-benchmark
N=${1:-2}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_round2.py -q -m gpu -k "two_gpu or peer_fedavg" 2>&1 | tail -2
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 scripts/p2p_bench.py 2>/dev/null | grep case > gpurun_out/r02_p2p_bench_${N}gpu.jsonl
cat gpurun_out/r02_p2p_bench_${N}gpu.jsonl
