#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_fedavg.py tests/test_gpu_compression.py tests/test_gpu_round.py -q -m gpu -x 2>&1 | tail -5
timeout 600 python scripts/fedavg_sweep.py --out gpurun_out/r02_fedavg_sweep_1gpu.json > /dev/null 2> gpurun_out/sweep.err; echo rc=$?
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02_fedavg_sweep_1gpu.json"))
for r in d.get("rows", d.get("fedavg", [])):
    print({k: r[k] for k in r if k in ("K", "P", "fedavg_ms", "frac_hbm", "q8_ms", "q8_frac_hbm", "GBs", "q8_GBs")})
PY
