#!/bin/bash
# Round 2, call 16: ROIAlign with compile-time footprint width + mapper (scan without atomics, batched staging, ballot voxel
# lists, 512-thread CTAs for large columns): tests, timings, launch lists, ncu --set full of the mapper's column / fuse kernels.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_mapper_gpu.py tests/test_pipeline_gpu.py tests/test_maskrcnn_gpu.py -m gpu -x -q > gpurun_out/r2l_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2l_pytest.log
tail -n 6 gpurun_out/r2l_pytest.log
for E in 1 8 32; do timeout 200 python tools/mapper_profile.py $E > gpurun_out/r2l_mapper_e$E.txt 2>&1; cat gpurun_out/r2l_mapper_e$E.txt | tail -n 2; done
for E in 1 32; do
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2l_mapper_launches_e$E.csv python tools/mapper_profile.py $E > /dev/null 2>&1
done
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"k_columns_warp|k_columns_cta|k_fuse|k_scan" -c 5 -f -o gpurun_out/r2l_mapper_e8 \
  python tools/mapper_profile.py 8 > gpurun_out/r2l_mapper_ncu.log 2>&1; echo "ncu mapper exit $?"
timeout 300 python tools/mrcnn_profile.py 8 bf16 > gpurun_out/r2l_ops_mrcnn_b8_bf16.txt 2>&1; grep -E "roi_align|^#" gpurun_out/r2l_ops_mrcnn_b8_bf16.txt
timeout 300 python tools/mrcnn_profile.py 32 bf16 > gpurun_out/r2l_ops_mrcnn_b32_bf16.txt 2>&1; grep -E "roi_align|^#" gpurun_out/r2l_ops_mrcnn_b32_bf16.txt
timeout 300 python tools/mrcnn_profile.py 1 tf32 > gpurun_out/r2l_ops_mrcnn_b1_tf32.txt 2>&1; grep -E "roi_align|^#" gpurun_out/r2l_ops_mrcnn_b1_tf32.txt
timeout 600 python bench.py --no-cpu-baseline --no-ref-gpu --no-latency --no-scaling-base --workload cfg2 > gpurun_out/r2l_cfg2.json 2> gpurun_out/r2l_cfg2.err; echo "bench exit $?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2l_cfg2.json"))
print("cfg2 value %.1f fps (%.2f ms) e2e %.1f fps conv frac %.3f glue %.3f ms clocks %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"], d["roofline"].get("glue_mapper_window_ms_per_step") or -1, d["clocks"]))
PY
