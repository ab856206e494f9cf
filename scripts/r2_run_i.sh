#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -30 > gpurun_out/r2i_tests.log
tail -4 gpurun_out/r2i_tests.log
for i in 1 2; do
timeout 300 python bench.py --no-extra --no-cpu-baseline > gpurun_out/r2i_bench_quick$i.json 2> /dev/null; echo "rc=$?"
done
timeout 300 python bench.py --no-extra --no-cpu-baseline --workload mnist_dp50 > gpurun_out/r2i_bench_m50.json 2> /dev/null; echo "rc=$?"
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2i_bench_*.json")):
    try:
        d = json.load(open(f))
        print(f, "round_ms", round(d["ms_per_step"], 4), "e2e_ms", round(d["e2e"]["ms_per_step"], 4), d["roofline"]["step_breakdown_ms"])
    except Exception as e:
        print(f, "unreadable", e)
PY
# per-role timeline of the resident conv kernels: needs the trace points compiled in (on this box only)
FLB_TRACE=1 python -m flb200.build --force > /dev/null 2>&1 && timeout 300 python scripts/conv_timeline.py > gpurun_out/r02_conv_timeline.txt 2> gpurun_out/conv_timeline.err; echo "timeline rc=$?"; wc -l gpurun_out/r02_conv_timeline.txt
