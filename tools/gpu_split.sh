#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_conv_gpu.py -m gpu -x -q > gpurun_out/pytest_conv.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_conv.log
tail -5 gpurun_out/pytest_conv.log
timeout 600 python tools/conv_split_sweep.py tf32 1 > gpurun_out/split_tf32_b1.txt 2>&1
timeout 600 python tools/conv_split_sweep.py bf16 1 > gpurun_out/split_bf16_b1.txt 2>&1
timeout 600 python tools/conv_split_sweep.py bf16 8 > gpurun_out/split_bf16_b8.txt 2>&1
cat gpurun_out/split_tf32_b1.txt
