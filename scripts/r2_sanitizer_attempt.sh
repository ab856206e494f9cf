#!/bin/bash
# round-2 batch C: tests, A/B of the fused kernels, ncu of the new kernels, compute-sanitizer
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -150 > gpurun_out/r2c_tests.log
tail -8 gpurun_out/r2c_tests.log
for cfg in "default:" "nofc1:FLB_NO_FUSED_FC1=1" "noadam:FLB_NO_FUSED_ADAM=1" "none:FLB_NO_FUSED_FC1=1 FLB_NO_FUSED_ADAM=1" "none_noswap:FLB_NO_FUSED_FC1=1 FLB_NO_FUSED_ADAM=1 FLB_FC_WGRAD_NO_SWAP=1"; do
  tag=${cfg%%:*}; envs=${cfg#*:}
  env $envs timeout 300 python bench.py --no-extra --no-cpu-baseline > gpurun_out/r2c_bench_$tag.json 2> gpurun_out/r2c_bench_$tag.err
  echo "bench $tag rc=$?"; tail -c 300 gpurun_out/r2c_bench_$tag.err
done
timeout 600 python bench.py > gpurun_out/r2c_bench_full.json 2> gpurun_out/r2c_bench_full.err
echo "full bench rc=$?"; tail -c 400 gpurun_out/r2c_bench_full.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2c_bench_*.json")):
    try:
        d = json.load(open(f))
        print(f, "round_ms", round(d["ms_per_step"], 4), "e2e_ms", round(d["e2e"]["ms_per_step"], 4), d["roofline"]["kernel"], d["roofline"]["step_breakdown_ms"])
        if "parity" in d: print(" parity", {k: v for k, v in d["parity"].items() if k.startswith("rel") or k == "pass"})
        for k, v in d.get("configs", {}).items():
            print(" ", k, {kk: vv for kk, vv in v.items() if kk in ("round_ms", "samples_per_s", "error")} if k != "fedavg_sweep" else [(r.get("K"), r.get("P"), r.get("frac_hbm"), r.get("q8_frac_hbm")) for r in v.get("rows", [])])
    except Exception as e:
        print(f, "unreadable", e)
PY
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"fc1_fused_kernel|FcWgradSwapT" -s 20 -c 4 -o gpurun_out/r2c_fc1 -f \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/r2c_ncu2.log 2>&1
ls -la gpurun_out/r2c_fc1.ncu-rep
# compute-sanitizer: memcheck over the tensor-core / fused-kernel tests, racecheck over the SimpleCNN tensor-core tests
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_tc.py tests/test_gpu_round2.py tests/test_gpu_compression.py -q -m gpu -x -k "not two_gpu" > gpurun_out/r2c_memcheck.log 2>&1
echo "memcheck rc=$?"; tail -5 gpurun_out/r2c_memcheck.log
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_gpu_tc.py -q -m gpu -x > gpurun_out/r2c_racecheck.log 2>&1
echo "racecheck rc=$?"; tail -5 gpurun_out/r2c_racecheck.log
