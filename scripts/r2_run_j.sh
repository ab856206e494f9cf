#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -30 > gpurun_out/r2j_tests.log
tail -4 gpurun_out/r2j_tests.log
for i in 1 2; do
timeout 300 python bench.py --no-extra --no-cpu-baseline > gpurun_out/r2j_bench_quick$i.json 2> /dev/null; echo "rc=$?"
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2j_bench_*.json")):
    try:
        d = json.load(open(f))
        print(f, "round_ms", round(d["ms_per_step"], 4), "e2e_ms", round(d["e2e"]["ms_per_step"], 4), d["roofline"]["step_breakdown_ms"])
    except Exception as e:
        print(f, "unreadable", e)
PY
