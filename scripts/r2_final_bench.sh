#!/bin/bash
# the driver's own invocations, for profiles/: default bench line, reference arm, smoke
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/r02_bench_1gpu.json 2> gpurun_out/r02_bench_1gpu.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_reference_arm.json 2>/dev/null; echo "ref rc=$?"
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02_bench_1gpu.json").read().strip().splitlines()[-1])
print("round_ms", round(d["ms_per_step"], 4), "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "parity", d["parity"].get("pass"), "cpu", round(d["cpu_baseline"]["value"]))
for k, v in d.get("configs", {}).items():
    print(k, {kk: vv for kk, vv in v.items() if kk in ("round_ms", "samples_per_s")})
PY
