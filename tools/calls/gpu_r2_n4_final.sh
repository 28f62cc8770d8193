#!/bin/bash
# Round 2: N4 kernels after the four-cells-per-thread sample kernel: tests and timing.
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_map_dataset_gpu.py -m gpu -q 2>&1 | tail -n 6
timeout 100 python tools/map_dataset_profile.py 2>&1 | tail -n 4
