#!/bin/bash
# full ncu capture of one whole training step (K = 10) of the current build + the launch list: inputs of scripts/collect_profiles.py
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 450 --csv --log-file gpurun_out/r02_launches_raw.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/r02_ncu1.log 2>&1
timeout 1200 ncu --set full --import-source on --clock-control none -s 400 -c 14 -o gpurun_out/r02_step_full -f \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/r02_ncu2.log 2>&1
ls -la gpurun_out/r02_step_full.ncu-rep
