#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_maskrcnn_gpu.py -m gpu -x -q 2>&1 | tail -3
timeout 300 python tools/mrcnn_profile.py 8 bf16 > gpurun_out/ops_mrcnn_b8_bf16.txt 2>&1
timeout 300 python tools/mrcnn_profile.py 1 tf32 > gpurun_out/ops_mrcnn_b1_tf32.txt 2>&1
grep -E "roi_align|graph replay|detections|rpn_nms" gpurun_out/ops_mrcnn_b8_bf16.txt gpurun_out/ops_mrcnn_b1_tf32.txt
