#!/bin/bash
# Round 2, call 14: GPU tests at HEAD (per-ROI ROIAlign with shared row tables, warp / CTA voxel-column kernels, quantile shortcut),
# mapper and Mask-RCNN op timings, mapper launch list under ncu, cfg2 bench.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2j_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2j_pytest.log
tail -n 15 gpurun_out/r2j_pytest.log
for E in 1 8 32; do timeout 200 python tools/mapper_profile.py $E > gpurun_out/r2j_mapper_e$E.txt 2>&1; cat gpurun_out/r2j_mapper_e$E.txt | tail -n 2; done
for E in 1 8; do
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2j_mapper_launches_e$E.csv python tools/mapper_profile.py $E > /dev/null 2>&1
done
timeout 300 python tools/mrcnn_profile.py 8 bf16 > gpurun_out/r2j_ops_mrcnn_b8_bf16.txt 2>&1; grep -E "roi_align|^#" gpurun_out/r2j_ops_mrcnn_b8_bf16.txt
timeout 300 python tools/mrcnn_profile.py 32 bf16 > gpurun_out/r2j_ops_mrcnn_b32_bf16.txt 2>&1; grep -E "roi_align|^#" gpurun_out/r2j_ops_mrcnn_b32_bf16.txt
timeout 600 python bench.py --no-cpu-baseline --no-ref-gpu --no-latency --no-scaling-base --workload cfg2 > gpurun_out/r2j_cfg2.json 2> gpurun_out/r2j_cfg2.err; echo "bench exit $?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2j_cfg2.json"))
print("cfg2 value %.1f fps (%.2f ms) e2e %.1f fps conv frac %.3f glue %.3f ms clocks %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"], d["roofline"].get("glue_mapper_window_ms_per_step") or -1, d["clocks"]))
PY
