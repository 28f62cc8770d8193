#!/bin/bash
# Round 2, last verification of HEAD on one GPU: full GPU tests, smoke, strict-parity margins, N4 kernel timing, default bench line.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_final2_pytest.log 2>&1; echo "pytest exit $?"; tail -n 3 gpurun_out/r02_final2_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 3
timeout 300 python tools/fp32_parity_probe.py prednet maskrcnn > gpurun_out/r02_fp32_probe2.txt 2>&1; cat gpurun_out/r02_fp32_probe2.txt
timeout 200 python tools/map_dataset_profile.py > gpurun_out/r02_map_dataset.txt 2>&1; cat gpurun_out/r02_map_dataset.txt
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 3 > gpurun_out/r02_final2_bench.json 2> gpurun_out/r02_final2_bench.err; echo "bench exit $?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/r02_final2_bench.json"))
print("cfg2 value %.1f fps (%.2f ms) e2e %.1f fps conv frac %.3f backbone %.3f clocks %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"], d["roofline"]["backbone"]["frac"], d["clocks"]))
print("latency", d["latency"]["frames_per_s"], d["latency"]["e2e_frames_per_s"], "cfg3", d["cfg3"]["value"], d["cfg3"]["e2e_value"])
print("others", d["roofline"]["largest_other_launches_ms"])
PY
