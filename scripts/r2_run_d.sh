#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/mma_microbench.py > gpurun_out/mma_microbench.jsonl 2> gpurun_out/mma_microbench.err
echo "microbench rc=$?"; grep UNROLLED gpurun_out/mma_microbench.jsonl | cut -c1-260; tail -3 gpurun_out/mma_microbench.err
