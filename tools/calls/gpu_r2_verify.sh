#!/bin/bash
# Round 2: HEAD verification as the driver will run it (GPU tests, smoke, both bench arms at N = 1).
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_verify_pytest.log 2>&1; echo "pytest exit $?"; tail -n 2 gpurun_out/r02_verify_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 4
timeout 600 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > gpurun_out/r02_verify_ref.json 2>/dev/null; echo "ref exit $?"; cut -c1-160 gpurun_out/r02_verify_ref.json
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 3 > gpurun_out/r02_verify_bench.json 2> gpurun_out/r02_verify_bench.err; echo "bench exit $?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/r02_verify_bench.json")); r=json.load(open("gpurun_out/r02_verify_ref.json"))
print("cfg2 value %.1f fps (%.2f ms) e2e %.1f fps conv frac %.3f clocks %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"], d["clocks"]))
print("same config:", d["config"] == r["config"], "e2e ratio %.1f" % (d["e2e"]["value"] / r["e2e"]["value"]))
print("keys", sorted(d.keys()))
PY
