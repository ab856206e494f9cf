#!/bin/bash
# The measurement batch behind profiles/r01_*: run on a 1-GPU box from the repo root (results land in gpurun_out/).
for i in 1 2; do timeout 600 python -m pytest tests -q -m gpu 2>&1 | tail -1; done
python bench.py > gpurun_out/f_bench1.json 2>gpurun_out/f_bench1.err
python bench.py --workload cifar_dp_q8 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/f_cifar.json 2>/dev/null
python bench.py --dp-mode per_sample --no-cpu-baseline > gpurun_out/f_ps.json 2>/dev/null
python bench.py --dp-mode none --no-cpu-baseline > gpurun_out/f_nodp.json 2>/dev/null
python bench.py --workload mnist_dp50 --no-cpu-baseline > gpurun_out/f_m50.json 2>/dev/null
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/f_ref.json 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 450 --csv --log-file gpurun_out/f_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/f_ncu1.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 300 --csv --log-file gpurun_out/f_cifar_launches.csv \
    python bench.py --workload cifar_dp_q8 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/f_ncu3.log 2>&1
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
