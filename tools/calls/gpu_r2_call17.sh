#!/bin/bash
# Round 2, call 17: full GPU test suite, mapper timings after the scan / fuse / staging changes, default bench line (N = 1).
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2m_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2m_pytest.log
tail -n 6 gpurun_out/r2m_pytest.log
for E in 1 8 32; do timeout 200 python tools/mapper_profile.py $E > gpurun_out/r2m_mapper_e$E.txt 2>&1; cat gpurun_out/r2m_mapper_e$E.txt | tail -n 2; done
for E in 1 32; do
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2m_mapper_launches_e$E.csv python tools/mapper_profile.py $E > /dev/null 2>&1
done
timeout 900 python bench.py --verbose > gpurun_out/r2m_bench.json 2> gpurun_out/r2m_bench.err; echo "bench exit $?"
tail -n 3 gpurun_out/r2m_bench.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2m_bench.json"))
print("cfg2 value %.1f fps (%.2f ms) e2e %.1f fps conv frac %.3f glue %.3f ms clocks %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"], d["roofline"].get("glue_mapper_window_ms_per_step") or -1, d["clocks"]))
print("latency", d["latency"]["frames_per_s"], d["latency"]["e2e_frames_per_s"], "cfg3", d["cfg3"]["value"], d["cfg3"]["e2e_value"])
print("ref_gpu", d.get("ref_gpu_eager")); print("cpu", d.get("cpu_baseline"))
PY
