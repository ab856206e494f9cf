#!/bin/bash
# round-2 batch A: GPU tests (with and without the fused classifier kernel) + bench lines (results land in gpurun_out/)
mkdir -p gpurun_out
echo "== tests, fused fc1 kernel OFF =="
FLB_NO_FUSED_FC1=1 timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -40 > gpurun_out/r2a_tests_nofc1.log
tail -12 gpurun_out/r2a_tests_nofc1.log
echo "== tests touching the SimpleCNN tensor-core path, fused fc1 kernel ON =="
timeout 900 python -m pytest tests/test_gpu_tc.py tests/test_gpu_training.py tests/test_gpu_round.py tests/test_gpu_round2.py -q -m gpu 2>&1 | tail -40 > gpurun_out/r2a_tests_fc1.log
tail -12 gpurun_out/r2a_tests_fc1.log
for cfg in "default:" "nofc1:FLB_NO_FUSED_FC1=1" "noadam:FLB_NO_FUSED_ADAM=1" "none:FLB_NO_FUSED_FC1=1 FLB_NO_FUSED_ADAM=1"; do
  tag=${cfg%%:*}; envs=${cfg#*:}
  env $envs timeout 300 python bench.py --no-extra --no-cpu-baseline > gpurun_out/r2a_bench_$tag.json 2> gpurun_out/r2a_bench_$tag.err
  echo "bench $tag rc=$?"; tail -c 300 gpurun_out/r2a_bench_$tag.err
done
FLB_NO_FUSED_FC1=1 timeout 600 python bench.py > gpurun_out/r2a_bench_full.json 2> gpurun_out/r2a_bench_full.err
echo "full bench rc=$?"; tail -c 400 gpurun_out/r2a_bench_full.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2a_ref.json 2>/dev/null
FLB_NO_FUSED_FC1=1 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok (fc1 unfused)')" 2>&1 | tail -1
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2a_bench_*.json")):
    try:
        d = json.load(open(f))
        print(f, "round_ms", round(d["ms_per_step"], 4), "e2e_ms", round(d["e2e"]["ms_per_step"], 4), d["roofline"]["kernel"], d["roofline"]["step_breakdown_ms"])
        if "parity" in d: print(" parity", d["parity"])
        for k, v in d.get("configs", {}).items():
            print(" ", k, {kk: vv for kk, vv in v.items() if kk in ("round_ms", "samples_per_s", "error")} if k != "fedavg_sweep" else v.get("rows", v))
        if "cpu_baseline" in d: print(" cpu", d["cpu_baseline"]["value"], d["cpu_baseline"]["kind"], d["cpu_baseline"].get("one_thread", {}).get("value"))
    except Exception as e:
        print(f, "unreadable", e)
PY
