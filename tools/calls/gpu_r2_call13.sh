#!/bin/bash
# Round 2, call 13: GPU tests at HEAD (micro-batched pipeline, global goal), cfg2 with 1 / 2 / 4 micro-batches back to back.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2i_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2i_pytest.log
tail -n 12 gpurun_out/r2i_pytest.log
F="--no-cpu-baseline --no-ref-gpu --no-latency --no-scaling-base --workload cfg2"
for mb in 1 2 4 1 2; do
  timeout 300 python bench.py $F --micro-batches $mb > gpurun_out/r2i_cfg2_mb$mb.json 2> gpurun_out/r2i_cfg2_mb$mb.err; echo "mb=$mb exit $?"
  python - <<EOF
import json
d=json.load(open("gpurun_out/r2i_cfg2_mb$mb.json"))
print("mb=$mb value %.1f fps (%.2f ms) e2e %.1f fps conv frac %.3f clocks %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"], d["clocks"]))
EOF
done
for mb in 1 2; do
  timeout 300 python bench.py --no-cpu-baseline --no-ref-gpu --no-latency --no-scaling-base --workload cfg3 --micro-batches $mb > gpurun_out/r2i_cfg3_mb$mb.json 2> gpurun_out/r2i_cfg3_mb$mb.err
  python - <<EOF
import json
d=json.load(open("gpurun_out/r2i_cfg3_mb$mb.json"))
print("cfg3 mb=$mb value %.1f fps (%.2f ms) e2e %.1f fps" % (d["value"], d["ms_per_step"], d["e2e"]["value"]))
EOF
done
