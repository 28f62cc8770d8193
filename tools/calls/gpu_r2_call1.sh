#!/bin/bash
# Round 2, call 1: reproduce the cfg3 stall with each opt-in autotune mode (watchdog names the configuration),
# baseline benches of cfg3 / cfg2 at the default modes, measured TF32 peak.
mkdir -p gpurun_out
for extra in 2 1; do
  echo "== PN_CONV_TUNE_EXTRA=$extra"
  PN_CONV_TUNE_EXTRA=$extra PN_CONV_TUNE_LOG=1 timeout 200 python bench.py --workload cfg3 --no-cpu-baseline --steps 5 \
      > gpurun_out/bisect_extra$extra.json 2> gpurun_out/bisect_extra$extra.err
  echo "exit $?"; tail -n 4 gpurun_out/bisect_extra$extra.err; cut -c1-300 gpurun_out/bisect_extra$extra.json
done
timeout 120 python tools/measure_tf32_peak.py gpurun_out/tf32_peak.json; cat gpurun_out/tf32_peak.json
timeout 300 python bench.py --workload cfg3 --no-cpu-baseline --steps 20 > gpurun_out/r2a_cfg3.json 2> gpurun_out/r2a_cfg3.err; echo "cfg3 exit $?"
timeout 300 python bench.py --workload cfg2 --no-cpu-baseline --steps 10 > gpurun_out/r2a_cfg2.json 2> gpurun_out/r2a_cfg2.err; echo "cfg2 exit $?"
cut -c1-400 gpurun_out/r2a_cfg3.json gpurun_out/r2a_cfg2.json
