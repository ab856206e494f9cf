#!/bin/bash
# usage: r2_run_r.sh N -- e2e A/B of the two collectives, alternating order
N=$1
mkdir -p gpurun_out
for mode in nccl p2p nccl p2p; do
  if [ $mode = nccl ]; then export FLB_NO_P2P=1; else unset FLB_NO_P2P; fi
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus $N --steps 20 --warmup 3 --no-extra > gpurun_out/r2r_${N}gpu_$mode.json 2> /dev/null
  python - <<PY
import json
d = json.load(open("gpurun_out/r2r_${N}gpu_$mode.json"))
print("$mode ${N} gpu round_ms", round(d["ms_per_step"], 4), "e2e_ms", round(d["e2e"]["ms_per_step"], 4), "wall", round(d["e2e"]["wall_ms_per_step"], 4))
PY
done
