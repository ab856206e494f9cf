#!/bin/bash
# usage: r2_ab.sh ENVVAR [tests]  -- GPU parity tests (optional), then the headline bench with ENVVAR unset / set, twice each
V=$1
mkdir -p gpurun_out
if [ "$2" = "tests" ]; then
timeout 1200 python -m pytest tests -q -m gpu -x 2>&1 | tail -4
fi
for rep in 1 2; do
for v in on off; do
if [ $v = off ]; then export $V=1; else unset $V; fi
timeout 300 python bench.py --no-extra --no-cpu-baseline > gpurun_out/r2ab_$v.json 2>/dev/null
python - <<PY
import json
d = json.loads(open("gpurun_out/r2ab_$v.json").read().strip().splitlines()[-1])
print("$V unset" if "$v" == "on" else "$V=1", "round_ms", round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["ms_per_step"], 4), d["roofline"].get("step_breakdown_ms"))
PY
done
done
