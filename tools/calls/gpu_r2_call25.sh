#!/bin/bash
# Round 2: whole-chain strict-parity test and the fp32 batch-vs-single test (printed margins).
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_pipeline_gpu.py tests/test_maskrcnn_gpu.py -m gpu -q -s -k "whole_chain or batch_equals_single_fp32" > gpurun_out/r02_whole_chain.txt 2>&1; echo "pytest exit $?"; tail -n 25 gpurun_out/r02_whole_chain.txt
