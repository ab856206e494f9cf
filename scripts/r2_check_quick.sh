#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu -x 2>&1 | tail -4
timeout 300 python bench.py --no-extra --no-cpu-baseline > gpurun_out/r2s.json 2>/dev/null
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2s.json").read().strip().splitlines()[-1])
print("round_ms", round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["ms_per_step"], 4), "launches", d["gpu_launches"])
PY
