#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --import-source on --clock-control none -k regex:fedavg_q8 -s 2 -c 1 -o gpurun_out/r02_q8_full -f python scripts/q8_probe.py > gpurun_out/q8_ncu.log 2>&1
ls -la gpurun_out/r02_q8_full.ncu-rep
