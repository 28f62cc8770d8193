#!/bin/bash
# Round 2: compute-sanitizer over the kernels rewritten this round (mapper columns / scan / fuse, ROIAlign).
mkdir -p gpurun_out
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_mapper_gpu.py -m gpu -x -q -k "wall_near or room_cat15 or room0 or batched" > gpurun_out/r02_sanitize_mapper_memcheck.log 2>&1; echo "mapper memcheck exit $?"
tail -n 4 gpurun_out/r02_sanitize_mapper_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_mapper_gpu.py -m gpu -x -q -k "wall_near or room_cat15 or room0" > gpurun_out/r02_sanitize_mapper_racecheck.log 2>&1; echo "mapper racecheck exit $?"
tail -n 4 gpurun_out/r02_sanitize_mapper_racecheck.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_maskrcnn_gpu.py -m gpu -x -q -k "roi_align" > gpurun_out/r02_sanitize_roi_memcheck.log 2>&1; echo "roi memcheck exit $?"
tail -n 4 gpurun_out/r02_sanitize_roi_memcheck.log
