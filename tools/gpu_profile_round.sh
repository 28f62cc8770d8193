#!/bin/bash
# ncu evidence for profiles/: (1) launch list of the default bench command, (2) DRAM traffic of every conv launch of
# that command, (3) --set full captures of representative conv layers (stand-alone, pn_conv_bench).
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_r01.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-profile > gpurun_out/launches_bench.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:conv_umma -c 800 --csv \
  --log-file gpurun_out/conv_traffic_r01.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-profile > gpurun_out/traffic_bench.log 2>&1
for spec in "A.fpn_out2 bf16 8" "A.res4.conv3 bf16 8" "A.res4.conv2 tf32 1"; do
  set -- $spec
  tag=$(echo "$1_$2_b$3" | tr '.' '_')
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_umma -s 4 -c 1 -f -o gpurun_out/r01_conv_$tag \
    python tools/conv_one.py $1 $2 $3 5 > gpurun_out/r01_conv_$tag.log 2>&1
  tail -1 gpurun_out/r01_conv_$tag.log
done
ls -la gpurun_out | tail -12
