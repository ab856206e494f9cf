#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -30 > gpurun_out/r2h_tests.log
tail -4 gpurun_out/r2h_tests.log
for i in 1 2; do
timeout 300 python bench.py --no-extra --no-cpu-baseline > gpurun_out/r2h_bench_quick$i.json 2> /dev/null; echo "rc=$?"
timeout 300 python bench.py --no-extra --no-cpu-baseline --workload cifar_dp_q8 --steps 5 > gpurun_out/r2h_bench_cifar$i.json 2> /dev/null; echo "rc=$?"
FLB_FUSED_ADAM=1 timeout 300 python bench.py --no-extra --no-cpu-baseline --workload cifar_dp_q8 --steps 5 > gpurun_out/r2h_bench_cifar_adam$i.json 2> /dev/null; echo "rc=$?"
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2h_bench_*.json")):
    try:
        d = json.load(open(f))
        print(f, "round_ms", round(d["ms_per_step"], 4), "e2e_ms", round(d["e2e"]["ms_per_step"], 4), {k: v for k, v in d["roofline"]["step_breakdown_ms"].items() if "fc" in k or "opt" in k})
    except Exception as e:
        print(f, "unreadable", e)
PY
