#!/bin/bash
# Round 2: N4 device kernels (tests + timing) and the strict-parity tests.
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_map_dataset_gpu.py -m gpu -q 2>&1 | tail -n 15
timeout 200 python tools/map_dataset_profile.py > gpurun_out/r02_map_dataset.txt 2>&1; cat gpurun_out/r02_map_dataset.txt
timeout 400 python -m pytest tests/test_conv_gpu.py tests/test_prednet_gpu.py tests/test_maskrcnn_gpu.py -m gpu -q -k "test_conv_fp32 or test_fp32_vs_fp32_oracle or reference_frame_fp32" 2>&1 | tail -n 15
