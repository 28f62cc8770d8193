#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_conv_gpu.py -m gpu -x -q > gpurun_out/pytest_conv.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_conv.log
tail -5 gpurun_out/pytest_conv.log
PN_CONV_DBG=1 python tools/conv_one.py A.res4.conv3 bf16 8 5 2>&1 | head -5
for spec in "A.res4.conv3 bf16 8" "A.res2.conv3 bf16 8" "A.res4.conv2 bf16 8" "A.fpn_out2 bf16 8" "A.res4.conv1 bf16 8" "A.res3.conv3 bf16 8" "A.res2.conv1 bf16 8" "A.fpn_out2 bf16 32" "A.res4.conv2 bf16 32"  "A.res4.conv3 bf16 32" "A.res4.conv3 tf32 1" "A.res4.conv1 tf32 1" "A.res4.conv2 tf32 1" "A.res2.conv3 tf32 1" "A.res3.conv3 tf32 1" "A.mask_fcn tf32 1"; do
  python tools/conv_one.py $spec 20
done
