#!/bin/bash
mkdir -p gpurun_out
for spec in "A.res4.conv3 bf16 8" "A.res2.conv3 bf16 8"; do
  set -- $spec
  tag=$(echo "$1_$2_b$3" | tr '.' '_')
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_umma -s 4 -c 1 -f -o gpurun_out/ncu2_$tag \
    python tools/conv_one.py $1 $2 $3 5 > gpurun_out/ncu2_$tag.log 2>&1
  tail -1 gpurun_out/ncu2_$tag.log
done
