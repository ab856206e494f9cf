#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/conv_timeline.py > gpurun_out/conv_timeline.txt 2> gpurun_out/conv_timeline.err; echo rc=$?; tail -3 gpurun_out/conv_timeline.err; wc -l gpurun_out/conv_timeline.txt
