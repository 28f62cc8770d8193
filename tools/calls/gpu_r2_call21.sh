#!/bin/bash
# Round 2, call 21: pixel-pair stem (two output pixels per GEMM row, N = 128): Mask-RCNN parity tests, op timings, default bench
# with the launch-configuration table dumped (the new stem shapes are the only ones the committed table does not cover).
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_maskrcnn_gpu.py tests/test_pipeline_gpu.py tests/test_determinism_gpu.py -m gpu -x -q > gpurun_out/r2s_pytest.log 2>&1; echo "pytest exit $?"; tail -n 5 gpurun_out/r2s_pytest.log
PN_TUNING_DUMP=gpurun_out/tuning_r2s.txt timeout 900 python bench.py --no-cpu-baseline --no-ref-gpu > gpurun_out/r2s_bench.json 2> gpurun_out/r2s_bench.err; echo "bench exit $?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2s_bench.json"))
print("cfg2 value %.1f fps (%.2f ms) e2e %.1f fps conv frac %.3f backbone %.3f clocks %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"], d["roofline"]["backbone"]["frac"], d["clocks"]))
print("latency", d["latency"]["frames_per_s"], d["latency"]["e2e_frames_per_s"], "cfg3", d["cfg3"]["value"], d["cfg3"]["e2e_value"])
print("others", d["roofline"]["largest_other_launches_ms"])
PY
timeout 300 python tools/mrcnn_profile.py 32 bf16 > gpurun_out/r2s_ops_mrcnn_b32_bf16.txt 2>&1; grep -E "stem|pack|resize|maxpool|^#" gpurun_out/r2s_ops_mrcnn_b32_bf16.txt
timeout 300 python tools/mrcnn_profile.py 1 tf32 > gpurun_out/r2s_ops_mrcnn_b1_tf32.txt 2>&1; grep -E "stem|pack|resize|maxpool|^#" gpurun_out/r2s_ops_mrcnn_b1_tf32.txt
