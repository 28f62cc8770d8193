#!/bin/bash
# Round 2, call 18: vectorised fuse kernel - mapper / pipeline tests, mapper timings and launch lists, cfg2 + cfg1 bench.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_mapper_gpu.py tests/test_pipeline_gpu.py tests/test_agent_prediction_gpu.py -m gpu -x -q > gpurun_out/r2n_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2n_pytest.log
tail -n 6 gpurun_out/r2n_pytest.log
for E in 1 8 32; do timeout 200 python tools/mapper_profile.py $E > gpurun_out/r2n_mapper_e$E.txt 2>&1; cat gpurun_out/r2n_mapper_e$E.txt | tail -n 2; done
for E in 1 8 32; do
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2n_mapper_launches_e$E.csv python tools/mapper_profile.py $E > /dev/null 2>&1
done
for wl in cfg2 cfg1; do
timeout 600 python bench.py --no-cpu-baseline --no-ref-gpu --no-latency --no-scaling-base --workload $wl > gpurun_out/r2n_$wl.json 2> gpurun_out/r2n_$wl.err; echo "bench exit $?"
python - <<PY
import json
d=json.load(open("gpurun_out/r2n_$wl.json"))
print("$wl value %.1f fps (%.2f ms) e2e %.1f fps conv frac %.3f glue %.3f ms clocks %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"], d["roofline"].get("glue_mapper_window_ms_per_step") or -1, d["clocks"]))
PY
done
