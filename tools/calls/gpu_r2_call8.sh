#!/bin/bash
# Round 2, call 8: soak of the tensor-memory fix (no forward-ordering event, both extra tuner modes on: 24 runs), GPU tests,
# default bench + tuning tables.
mkdir -p gpurun_out
B="python bench.py --workload cfg3 --mode overlapped --no-cpu-baseline --no-ref-gpu --no-latency --no-profile --steps 5"
fails=0
for i in $(seq 1 24); do
  env PN_DEBUG_NO_FWD_ORDER=1 PN_CONV_TUNING_FILE=none timeout 30 $B > gpurun_out/p8_soak.json 2> gpurun_out/p8_soak.err || fails=$((fails+1))
done
echo "soak (alloc after wait, no event, extras on): $fails of 24 runs did not finish"
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2e_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2e_pytest.log
tail -n 25 gpurun_out/r2e_pytest.log
PN_CONV_TUNING_FILE=none PN_TUNING_DUMP=gpurun_out/tuning_cfg2_cfg1.txt timeout 600 python bench.py --verbose > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err; echo "bench exit $?"
grep -v "^\[tune\]" gpurun_out/r2e_bench.err | tail -n 4; cat gpurun_out/r2e_bench.json
PN_CONV_TUNING_FILE=none PN_TUNING_DUMP=gpurun_out/tuning_cfg3.txt timeout 300 python bench.py --workload cfg3 --no-cpu-baseline --no-ref-gpu --no-latency > gpurun_out/r2e_cfg3.json 2> gpurun_out/r2e_cfg3.err; echo "cfg3 exit $?"
cut -c1-600 gpurun_out/r2e_cfg3.json
