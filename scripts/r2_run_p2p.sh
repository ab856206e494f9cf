#!/bin/bash
# usage: r2_run_p2p.sh N [tests]  -- peer-memory FedAvg: 2-GPU parity tests (optional), micro-benchmark, bench line under torchrun
N=$1
mkdir -p gpurun_out
nvidia-smi -L | wc -l
if [ "$2" = "tests" ]; then
timeout 900 python -m pytest tests/test_gpu_round2.py -q -m gpu -x -k "two_gpu or peer_fedavg" 2>&1 | tail -15 > gpurun_out/r2_p2p_tests.log
tail -8 gpurun_out/r2_p2p_tests.log
fi
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 scripts/p2p_bench.py > gpurun_out/r02_p2p_bench_${N}gpu.jsonl 2> gpurun_out/r02_p2p_bench_${N}gpu.err
echo "p2p bench rc=$?"; cat gpurun_out/r02_p2p_bench_${N}gpu.jsonl; tail -c 600 gpurun_out/r02_p2p_bench_${N}gpu.err
for mode in p2p nccl; do
  if [ $mode = nccl ]; then export FLB_NO_P2P=1; else unset FLB_NO_P2P; fi
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus $N --steps 10 --warmup 3 --no-extra > gpurun_out/r02_bench_${N}gpu_$mode.json 2> gpurun_out/r02_bench_${N}gpu_$mode.err
  echo "bench $mode rc=$?"; tail -c 300 gpurun_out/r02_bench_${N}gpu_$mode.err
  python - <<PY
import json
d = json.load(open("gpurun_out/r02_bench_${N}gpu_$mode.json"))
print("$mode ${N} gpu round_ms", d["ms_per_step"], "value", d["value"], "e2e", d["e2e"]["value"])
PY
done
