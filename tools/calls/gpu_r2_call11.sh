#!/bin/bash
# Round 2, call 11: full GPU tests, smoke, per-op profiles (staged ROIAlign, voxel-column occupancy), default bench.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2g_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2g_pytest.log
tail -n 25 gpurun_out/r2g_pytest.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -n 6
timeout 300 python tools/mrcnn_profile.py 8 bf16 > gpurun_out/r2g_ops_mrcnn_b8_bf16.txt 2>&1; grep -E "roi_align|^#" gpurun_out/r2g_ops_mrcnn_b8_bf16.txt
timeout 300 python tools/mrcnn_profile.py 32 bf16 > gpurun_out/r2g_ops_mrcnn_b32_bf16.txt 2>&1; grep -E "roi_align|^#" gpurun_out/r2g_ops_mrcnn_b32_bf16.txt
timeout 300 python tools/mapper_profile.py 8 > gpurun_out/r2g_mapper_e8.txt 2>&1; tail -n 4 gpurun_out/r2g_mapper_e8.txt
timeout 300 python tools/mapper_profile.py 1 > gpurun_out/r2g_mapper_e1.txt 2>&1; tail -n 4 gpurun_out/r2g_mapper_e1.txt
timeout 900 python bench.py --verbose > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err; echo "bench exit $?"
tail -n 3 gpurun_out/r2g_bench.err; cat gpurun_out/r2g_bench.json
