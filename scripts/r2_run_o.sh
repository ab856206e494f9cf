#!/bin/bash
mkdir -p gpurun_out
timeout 300 python bench.py --no-extra --no-cpu-baseline --workload mnist_dp50 --steps 5 > gpurun_out/r2o_mnist50.json 2>/dev/null
timeout 300 python bench.py --no-extra --no-cpu-baseline --workload cifar_dp_q8 --steps 3 > gpurun_out/r2o_cifar.json 2>/dev/null
python - <<'PY'
import json
for f in ("gpurun_out/r2o_mnist50.json", "gpurun_out/r2o_cifar.json"):
    d = json.loads(open(f).read().strip().splitlines()[-1])
    print(f, "round_ms", d["ms_per_step"])
    bd = d["roofline"].get("step_breakdown_ms")
    print(sorted(bd.items(), key=lambda kv: -kv[1]))
    print("sum", sum(bd.values()))
PY
