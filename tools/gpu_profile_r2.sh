#!/bin/bash
# Round-2 ncu evidence for profiles/ (one GPU): steady-state launch lists of cfg2 (headline) and cfg1 (latency) with the
# committed tuning tables (no build-time launches), DRAM traffic of the conv launches without ncu's per-kernel L2 flush,
# --set full captures of representative conv layers at batch 32.
mkdir -p gpurun_out
B="--no-cpu-baseline --no-ref-gpu --no-latency --no-profile"
for wl in cfg2 cfg1; do
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r02_launches_$wl.csv \
    python bench.py --workload $wl --steps 2 --warmup 3 $B > gpurun_out/r02_launches_$wl.log 2>&1
  echo "launch list $wl: exit $? $(wc -l < gpurun_out/r02_launches_$wl.csv) lines"
  timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --cache-control none --clock-control none \
    -k regex:conv_umma -c 4000 --csv --log-file gpurun_out/r02_conv_traffic_$wl.csv \
    python bench.py --workload $wl --steps 1 --warmup 3 $B > gpurun_out/r02_traffic_$wl.log 2>&1
  echo "traffic $wl: exit $? $(wc -l < gpurun_out/r02_conv_traffic_$wl.csv) lines"
done
for spec in "A.res4.conv3 bf16 32" "A.res4.conv2 bf16 32" "A.fpn_out2 bf16 32" "A.res2.conv3 bf16 32"; do
  set -- $spec
  tag=$(echo "$1_$2_b$3" | tr '.' '_')
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_umma -s 4 -c 1 -f -o gpurun_out/r02_conv_$tag \
    python tools/conv_one.py $1 $2 $3 5 > gpurun_out/r02_conv_$tag.log 2>&1
  tail -1 gpurun_out/r02_conv_$tag.log
done
timeout 600 python -m pytest tests/test_global_goal_gpu.py tests/test_gather_gpu.py tests/test_determinism_gpu.py -x -q 2>&1 | tail -n 8
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_roi_align -c 1 -f -o gpurun_out/r02_roi_align_b8 \
  python tools/mrcnn_profile.py 8 bf16 > gpurun_out/r02_roi_align_b8.log 2>&1
ls -la gpurun_out | grep r02_ | tail -20
