#!/bin/bash
# ncu --set full on individual conv layers (one captured launch each, after 4 warm launches)
mkdir -p gpurun_out
for spec in "A.res4.conv3 bf16 8" "A.res2.conv3 bf16 8" "A.res4.conv2 bf16 8" "A.fpn_out2 bf16 8" "A.res4.conv1 bf16 8" "A.res4.conv2 tf32 1" "A.res4.conv3 tf32 1"; do
  set -- $spec
  tag=$(echo "$1_$2_b$3" | tr '.' '_')
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_umma -s 4 -c 1 -f -o gpurun_out/ncu_$tag \
    python tools/conv_one.py $1 $2 $3 5 > gpurun_out/ncu_$tag.log 2>&1
  tail -1 gpurun_out/ncu_$tag.log
done
for spec in "A.res4.conv3 bf16 8" "A.res2.conv3 bf16 8" "A.res4.conv2 bf16 8" "A.fpn_out2 bf16 8" "A.res4.conv1 bf16 8" "A.fpn_out2 bf16 32" "A.res4.conv2 bf16 32"  "A.res4.conv3 bf16 32"; do
  python tools/conv_one.py $spec 20
done
