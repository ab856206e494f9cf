#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tc.py tests/test_gpu_cifar.py tests/test_gpu_training.py tests/test_gpu_round.py -q -m gpu -x 2>&1 | tail -30 > gpurun_out/r2g_tests.log
tail -4 gpurun_out/r2g_tests.log
timeout 300 python scripts/conv_timeline.py > gpurun_out/conv_timeline.txt 2> gpurun_out/conv_timeline.err; echo timeline rc=$?
timeout 300 python bench.py --no-extra --no-cpu-baseline > gpurun_out/r2g_bench_quick.json 2> gpurun_out/r2g_bench_quick.err; echo "rc=$?"
timeout 300 python bench.py --no-extra --no-cpu-baseline --workload cifar_dp_q8 --steps 5 > gpurun_out/r2g_bench_cifar.json 2> gpurun_out/r2g_bench_cifar.err; echo "rc=$?"
python - <<'PY'
import json
for f in ("r2g_bench_quick", "r2g_bench_cifar"):
    try:
        d = json.load(open(f"gpurun_out/{f}.json"))
        print(f, "round_ms", round(d["ms_per_step"], 4), "e2e_ms", round(d["e2e"]["ms_per_step"], 4), d["roofline"]["kernel"], {k: v for k, v in d["roofline"]["step_breakdown_ms"].items() if "conv" in k})
    except Exception as e:
        print(f, "unreadable", e)
PY
