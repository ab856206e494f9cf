#!/usr/bin/env python
"""Copy one measurement batch (scripts/r2_profile.sh -> gpurun_out/r02_*) into profiles/ and derive the summaries:

    python scripts/collect_profiles.py

  * bench lines (json) as they are;
  * ncu launch lists -> per-kernel summaries (scripts/summarize_launches.py);
  * the `ncu --set full` capture of one training step -> one CSV row per launch (scripts/summarize_ncu_full.py) and
    profiles/ncu_traffic.json (dram__bytes_read.sum + dram__bytes_write.sum per launch, keyed by bench step label), which
    bench.py reports as roofline.traffic together with the commit the capture was made at.
"""
import csv
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
for f in ("r02_bench_1gpu.json", "r02_reference_arm.json", "r02_bench_1gpu_cifar100_q8.json", "r02_bench_1gpu_mnist50.json",
          "r02_bench_1gpu_per_sample_dp.json", "r02_bench_1gpu_cifar100_q8_per_sample.json", "r02_fedavg_sweep_1gpu.json", "r02_mma_microbench.jsonl", "r02_launches_raw.csv", "r02_conv_timeline.txt",
          "r02_bench_2gpu.json", "r02_bench_4gpu.json", "r02_bench_8gpu.json", "r02_2gpu_tests.log"):
    if os.path.exists(os.path.join(G, f)):
        shutil.copy(os.path.join(G, f), os.path.join(P, f))
py = sys.executable
for raw, out, title in (("r02_launches_raw.csv", "r02_launches_summary.csv",
                         "ncu launch list, round-2 final build (bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra; -s 300 -c 450)"),
                        ("r02_cifar_launches_raw.csv", "r02_cifar_launches_summary.csv",
                         "ncu launch list, CIFAR10CNN 100 clients q8 (bench.py --workload cifar_dp_q8 --steps 1 --warmup 3; -s 200 -c 300)")):
    if os.path.exists(os.path.join(G, raw)):
        txt = subprocess.run([py, os.path.join(ROOT, "scripts", "summarize_launches.py"), os.path.join(G, raw), title],
                             capture_output=True, text=True).stdout
        open(os.path.join(P, out), "w").write("# per-launch times are cold-cache and serialised: compare SHARES\n" + txt)
rep = os.path.join(G, "r02_step_full.ncu-rep")
if os.path.exists(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    open("/tmp/r02raw.csv", "w").write(raw)
    title = ("ncu --set full --clock-control none, bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extra (K = 10 clients, TF32 path, "
             "round-2 final build); one row per captured launch (-s 400 -c 14)")
    txt = subprocess.run([py, os.path.join(ROOT, "scripts", "summarize_ncu_full.py"), "/tmp/r02raw.csv", title], capture_output=True, text=True).stdout
    open(os.path.join(P, "r02_step_ncu_full_summary.csv"), "w").write(txt)
    rows = list(csv.reader(raw.splitlines()))
    hdr = rows[0]
    ki, ri, wi = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
    units = rows[1]
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    label = [("conv1_fwd_pool_kernel", "conv1_fwd_pool"), ("ConvFwdHaloT<32, 64, 1>", "conv2_fwd_pool"), ("fc1_fused_kernel", "fc1_fused"),
             ("head_wgrad_kernel", "head_wgrad"), ("FcWgradSwapT<3136, 0>", "fc1_wgrad"), ("unpool2_kernel", "unpool2"),
             ("ConvWgradHaloT<32, 64, 3, 0>", "conv2_wgrad"), ("ConvDgradHaloT<32, 64>", "conv2_dgrad"), ("conv1_bwd", "conv1_wgrad"),
             ("optimizer_kernel", "optimizer")]
    traffic = {}
    for r in rows[2:]:
        for pat, lab in label:
            if pat in r[ki] and lab not in traffic:
                traffic[lab] = int(float(r[ri]) * scale.get(units[ri], 1.0) + float(r[wi]) * scale.get(units[wi], 1.0))
    tj_path = os.path.join(P, "ncu_traffic.json")
    tj = json.load(open(tj_path)) if os.path.exists(tj_path) else {}
    tj.setdefault("simple_cnn", {})["10"] = traffic
    head = subprocess.run(["git", "-C", ROOT, "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
    tj["_capture"] = f"ncu --set full of bench.py --steps 1 --warmup 3 (scripts/r2_profile.sh), build at commit {head}"
    json.dump(tj, open(tj_path, "w"), indent=1)
    print("traffic", traffic)
print("collected")
