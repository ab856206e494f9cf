#!/bin/bash
# round-2 profile batch (1 GPU): final bench lines + ncu evidence, all into gpurun_out/r02_*
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/r02_bench_1gpu.json 2> gpurun_out/r02_bench_1gpu.err; echo "bench rc=$?"; tail -c 300 gpurun_out/r02_bench_1gpu.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_reference_arm.json 2>/dev/null; echo "ref rc=$?"
timeout 300 python bench.py --no-extra --no-cpu-baseline --workload cifar_dp_q8 --steps 5 > gpurun_out/r02_bench_1gpu_cifar100_q8.json 2>/dev/null
timeout 300 python bench.py --no-extra --no-cpu-baseline --workload mnist_dp50 > gpurun_out/r02_bench_1gpu_mnist50.json 2>/dev/null
timeout 300 python bench.py --no-extra --no-cpu-baseline --dp-mode per_sample > gpurun_out/r02_bench_1gpu_per_sample_dp.json 2>/dev/null
timeout 300 python bench.py --no-extra --no-cpu-baseline --workload cifar_dp_q8 --dp-mode per_sample --steps 3 > gpurun_out/r02_bench_1gpu_cifar100_q8_per_sample.json 2>/dev/null
timeout 600 python scripts/fedavg_sweep.py --out gpurun_out/r02_fedavg_sweep_1gpu.json > /dev/null 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:fedavg_q8 -s 2 -c 1 -o gpurun_out/r02_q8_full -f python scripts/q8_probe.py > gpurun_out/q8_ncu.log 2>&1
timeout 300 python scripts/mma_microbench.py > gpurun_out/r02_mma_microbench.jsonl 2>/dev/null
# ncu launch list of the bench command (per-launch times cold-cache and serialised: compare SHARES)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 450 --csv --log-file gpurun_out/r02_launches_raw.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/r02_ncu1.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 300 --csv --log-file gpurun_out/r02_cifar_launches_raw.csv \
    python bench.py --workload cifar_dp_q8 --steps 1 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/r02_ncu3.log 2>&1
# full capture of one whole training step (K = 10): DRAM traffic, tensor-pipe activity per kernel
timeout 1200 ncu --set full --import-source on --clock-control none -s 400 -c 14 -o gpurun_out/r02_step_full -f \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/r02_ncu2.log 2>&1
ls -la gpurun_out/r02_step_full.ncu-rep
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
