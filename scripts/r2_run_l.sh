#!/bin/bash
# CIFAR10CNN per-sample DP-SGD: parity tests
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_cifar.py tests/test_gpu_tc.py tests/test_gpu_training.py -q -m gpu -x -k "per_sample or dp" 2>&1 | tail -40 > gpurun_out/r2l_tests.log
cat gpurun_out/r2l_tests.log
