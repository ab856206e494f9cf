#!/bin/bash
# 2 GPUs: the sharded-round parity test, then the bench line under torchrun
mkdir -p gpurun_out
nvidia-smi -L | head -4
timeout 900 python -m pytest tests/test_gpu_round2.py tests/test_gpu_round.py -q -m gpu -k "two_gpu or round_sgd" 2>&1 | tail -15 > gpurun_out/r2_2gpu_tests.log
tail -6 gpurun_out/r2_2gpu_tests.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2_bench_2gpu.json 2> gpurun_out/r2_bench_2gpu.err
echo "bench rc=$?"; tail -c 400 gpurun_out/r2_bench_2gpu.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2_bench_2gpu.json"))
print("2gpu round_ms", d["ms_per_step"], "value", d["value"], "e2e", d["e2e"]["value"])
for k, v in d.get("configs", {}).items():
    print(" ", k, {kk: vv for kk, vv in v.items() if kk in ("round_ms", "samples_per_s", "error")})
PY
