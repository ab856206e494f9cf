#!/bin/bash
# re-entry sanity run: GPU test suite + default bench on the restored checkpoint
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -30 > gpurun_out/r2check_tests.log
tail -4 gpurun_out/r2check_tests.log
timeout 600 python bench.py > gpurun_out/r2check_bench.json 2> gpurun_out/r2check_bench.err; echo "rc=$?"
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r2check_bench.json").read().strip().splitlines()[-1])
    print("round_ms", round(d["ms_per_step"], 4), "e2e_ms", round(d["e2e"]["ms_per_step"], 4), d["roofline"].get("step_breakdown_ms"))
    for k, v in d.get("configs", {}).items():
        print(k, {kk: vv for kk, vv in v.items() if kk in ("round_ms", "samples_per_s")})
except Exception as e:
    print("unreadable", e)
PY
