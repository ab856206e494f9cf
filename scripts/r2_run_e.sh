#!/bin/bash
# round-2 batch E: tests + benches after the persistent-kernel fixes
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -60 > gpurun_out/r2e_tests.log
tail -6 gpurun_out/r2e_tests.log
timeout 300 python bench.py --no-extra --no-cpu-baseline > gpurun_out/r2e_bench_quick.json 2> gpurun_out/r2e_bench_quick.err; echo "rc=$?"
timeout 300 python bench.py --no-extra --no-cpu-baseline --workload cifar_dp_q8 --steps 5 > gpurun_out/r2e_bench_cifar.json 2> gpurun_out/r2e_bench_cifar.err; echo "rc=$?"
python - <<'PY'
import json
for f in ("r2e_bench_quick", "r2e_bench_cifar"):
    try:
        d = json.load(open(f"gpurun_out/{f}.json"))
        print(f, "round_ms", round(d["ms_per_step"], 4), "e2e_ms", round(d["e2e"]["ms_per_step"], 4), d["roofline"]["kernel"], d["roofline"]["step_breakdown_ms"])
    except Exception as e:
        print(f, "unreadable", e)
PY
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"conv_resident_kernel" -s 8 -c 4 -o gpurun_out/r2e_conv -f \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/r2e_ncu.log 2>&1
ls -la gpurun_out/r2e_conv.ncu-rep
