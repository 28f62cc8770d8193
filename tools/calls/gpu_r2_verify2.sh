#!/bin/bash
# Round 2: last verification of HEAD on one GPU: full GPU tests, smoke, default bench line, op timings, launch list of cfg2.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_verify_pytest.log 2>&1; echo "pytest exit $?"; tail -n 2 gpurun_out/r02_verify_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 4
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 3 > gpurun_out/r02_verify_bench.json 2> gpurun_out/r02_verify_bench.err; echo "bench exit $?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/r02_verify_bench.json"))
print("cfg2 value %.1f fps (%.2f ms) e2e %.1f fps conv frac %.3f backbone %.3f clocks %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"], d["roofline"]["backbone"]["frac"], d["clocks"]))
print("latency", d["latency"]["frames_per_s"], d["latency"]["e2e_frames_per_s"], "cfg3", d["cfg3"]["value"], d["cfg3"]["e2e_value"])
print("others", d["roofline"]["largest_other_launches_ms"])
PY
timeout 300 python tools/mrcnn_profile.py 32 bf16 > gpurun_out/r02_final_ops_mrcnn_b32_bf16.txt 2>&1; tail -n 1 gpurun_out/r02_final_ops_mrcnn_b32_bf16.txt
B="--no-cpu-baseline --no-ref-gpu --no-latency --no-profile --no-scaling-base"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2400 --csv --log-file gpurun_out/r02_launches_cfg2.csv \
    python bench.py --workload cfg2 --steps 2 --warmup 3 $B > gpurun_out/r02_launches_cfg2.log 2>&1
echo "launch list cfg2: exit $? $(wc -l < gpurun_out/r02_launches_cfg2.csv) lines"
