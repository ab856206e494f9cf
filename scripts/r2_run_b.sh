#!/bin/bash
# round-2 batch B: fixed tests, A/B of the fused kernels (graph-replayed round time), ncu of the two new kernels
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -150 > gpurun_out/r2b_tests.log
tail -8 gpurun_out/r2b_tests.log
for cfg in "default:" "nofc1:FLB_NO_FUSED_FC1=1" "noadam:FLB_NO_FUSED_ADAM=1" "none:FLB_NO_FUSED_FC1=1 FLB_NO_FUSED_ADAM=1" "nocoop:FLB_FC1_NO_COOP=1" "nocoop_noadam:FLB_FC1_NO_COOP=1 FLB_NO_FUSED_ADAM=1"; do
  tag=${cfg%%:*}; envs=${cfg#*:}
  env $envs timeout 300 python bench.py --no-extra --no-cpu-baseline > gpurun_out/r2b_bench_$tag.json 2> gpurun_out/r2b_bench_$tag.err
  echo "bench $tag rc=$?"; tail -c 300 gpurun_out/r2b_bench_$tag.err
done
timeout 600 python bench.py > gpurun_out/r2b_bench_full.json 2> gpurun_out/r2b_bench_full.err
echo "full bench rc=$?"; tail -c 400 gpurun_out/r2b_bench_full.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2b_bench_*.json")):
    try:
        d = json.load(open(f))
        print(f, "round_ms", round(d["ms_per_step"], 4), "e2e_ms", round(d["e2e"]["ms_per_step"], 4), d["roofline"]["kernel"], d["roofline"]["step_breakdown_ms"])
        if "parity" in d: print(" parity", d["parity"])
        for k, v in d.get("configs", {}).items():
            print(" ", k, {kk: vv for kk, vv in v.items() if kk in ("round_ms", "samples_per_s", "error")} if k != "fedavg_sweep" else "")
    except Exception as e:
        print(f, "unreadable", e)
PY
# ncu: launch list of the default configuration, then full captures of the two new kernels
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 450 --csv --log-file gpurun_out/r2b_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/r2b_ncu1.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"fc1_fused_kernel|FcWgradT" -s 40 -c 4 -o gpurun_out/r2b_fc1 -f \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/r2b_ncu2.log 2>&1
ls -la gpurun_out/r2b_fc1.ncu-rep
