#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_conv_gpu.py -m gpu -x -q 2>&1 | tail -6
for spec in "A.fpn_out2 bf16 8" "A.res4.conv2 bf16 8" "A.res4.conv3 bf16 8" "A.res4.conv1 bf16 8" "A.res2.conv3 bf16 8" "A.fpn_out2 bf16 32" "A.res4.conv2 bf16 32" "A.res4.conv3 bf16 32" "A.fpn_out2 tf32 1" "A.mask_fcn tf32 1" "A.res4.conv2 tf32 1" "A.fpn_out3 tf32 1" "C.psp.bottleneck bf16 8"; do
  a=$(python tools/conv_one.py $spec 20 0x4000 | tail -1)   # pairs forbidden
  b=$(python tools/conv_one.py $spec 20 0x2100 2>&1 | tail -1)   # pair forced, N tile 256
  c=$(python tools/conv_one.py $spec 20 0 | tail -1)        # automatic
  echo "$a || PAIR256: $b || AUTO: $c"
done
