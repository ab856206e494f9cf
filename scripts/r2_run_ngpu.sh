#!/bin/bash
# usage: r2_run_ngpu.sh N   -- the bench line under torchrun on N GPUs (weak scaling headline + sub-configs)
N=$1
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r02_bench_${N}gpu.json 2> gpurun_out/r02_bench_${N}gpu.err
echo "bench rc=$?"; tail -c 300 gpurun_out/r02_bench_${N}gpu.err
python - <<PY
import json
d = json.load(open("gpurun_out/r02_bench_${N}gpu.json"))
print("${N} gpu round_ms", d["ms_per_step"], "value", d["value"], "e2e", d["e2e"]["value"])
for k, v in d.get("configs", {}).items():
    print(" ", k, {kk: vv for kk, vv in v.items() if kk in ("round_ms", "samples_per_s", "error")} if k != "fedavg_sweep" else [(r.get("K"), r.get("P"), r.get("fedavg_ms"), r.get("frac_hbm")) for r in v.get("rows", [])])
PY
