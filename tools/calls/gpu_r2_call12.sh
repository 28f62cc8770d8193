#!/bin/bash
# Round 2, call 12: GPU tests at HEAD, ROIAlign timing (restructured per-bin loop), default bench.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2h_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2h_pytest.log
tail -n 12 gpurun_out/r2h_pytest.log
timeout 300 python tools/mrcnn_profile.py 8 bf16 > gpurun_out/r2h_ops_mrcnn_b8_bf16.txt 2>&1; grep -E "roi_align|^#" gpurun_out/r2h_ops_mrcnn_b8_bf16.txt
timeout 300 python tools/mrcnn_profile.py 32 bf16 > gpurun_out/r2h_ops_mrcnn_b32_bf16.txt 2>&1; grep -E "roi_align|^#" gpurun_out/r2h_ops_mrcnn_b32_bf16.txt
timeout 900 python bench.py --verbose > gpurun_out/r2h_bench.json 2> gpurun_out/r2h_bench.err; echo "bench exit $?"
tail -n 3 gpurun_out/r2h_bench.err; cat gpurun_out/r2h_bench.json
