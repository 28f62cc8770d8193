#!/bin/bash
# Round 2, final evidence pass (one GPU): full GPU test suite, the default bench line, steady-state launch lists of cfg2 / cfg1,
# DRAM traffic of the conv launches, ncu --set full of the final ROIAlign and mapper kernels, mapper timings.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_final_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02_final_pytest.log
tail -n 4 gpurun_out/r02_final_pytest.log
timeout 900 python bench.py > gpurun_out/r02_final_bench.json 2> gpurun_out/r02_final_bench.err; echo "bench exit $?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/r02_final_bench.json"))
print("cfg2 value %.1f fps (%.2f ms) e2e %.1f fps conv frac %.3f backbone %.3f glue %.3f ms clocks %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"], d["roofline"]["backbone"]["frac"], d["roofline"].get("glue_mapper_window_ms_per_step") or -1, d["clocks"]))
print("latency", d["latency"]["frames_per_s"], d["latency"]["e2e_frames_per_s"], "cfg3", d["cfg3"]["value"], d["cfg3"]["e2e_value"])
print("others", d["roofline"]["largest_other_launches_ms"])
PY
B="--no-cpu-baseline --no-ref-gpu --no-latency --no-profile --no-scaling-base"
for wl in cfg2 cfg1; do
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r02_launches_$wl.csv \
    python bench.py --workload $wl --steps 2 --warmup 3 $B > gpurun_out/r02_launches_$wl.log 2>&1
  echo "launch list $wl: exit $? $(wc -l < gpurun_out/r02_launches_$wl.csv) lines"
  timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --cache-control none --clock-control none \
    -k regex:conv_umma -c 4000 --csv --log-file gpurun_out/r02_conv_traffic_$wl.csv \
    python bench.py --workload $wl --steps 1 --warmup 3 $B > gpurun_out/r02_traffic_$wl.log 2>&1
  echo "traffic $wl: exit $? $(wc -l < gpurun_out/r02_conv_traffic_$wl.csv) lines"
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_roi_align -c 1 -f -o gpurun_out/r02_roi_align_final_b8 \
  python tools/mrcnn_profile.py 8 bf16 > gpurun_out/r02_roi_align_final_b8.log 2>&1; echo "ncu roi exit $?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"k_columns_warp|k_columns_cta|k_fuse|k_scan" -c 5 -f -o gpurun_out/r02_mapper_final_e8 \
  python tools/mapper_profile.py 8 > gpurun_out/r02_mapper_final_ncu.log 2>&1; echo "ncu mapper exit $?"
for E in 1 8 32; do timeout 200 python tools/mapper_profile.py $E > gpurun_out/r02_mapper_final_e$E.txt 2>&1; tail -n 1 gpurun_out/r02_mapper_final_e$E.txt
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_mapper_final_launches_e$E.csv python tools/mapper_profile.py $E > /dev/null 2>&1
done
timeout 300 python tools/mrcnn_profile.py 32 bf16 > gpurun_out/r02_final_ops_mrcnn_b32_bf16.txt 2>&1; tail -n 1 gpurun_out/r02_final_ops_mrcnn_b32_bf16.txt
timeout 300 python tools/prednet_profile.py 32 24 240 bf16 > gpurun_out/r02_final_ops_prednet_b32_bf16.txt 2>&1; tail -n 1 gpurun_out/r02_final_ops_prednet_b32_bf16.txt
