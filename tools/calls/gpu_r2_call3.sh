#!/bin/bash
# Round 2, call 3: GPU tests at HEAD, locate the two-CTAs-per-SM stall inside bench.py (verbose), default bench line,
# launch-configuration tables of the three workloads.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2c_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2c_pytest.log
tail -n 5 gpurun_out/r2c_pytest.log
echo "== extra=2 cfg3 verbose"
PN_CONV_TUNE_EXTRA=2 PN_CONV_TUNING_FILE=none timeout 100 python bench.py --workload cfg3 --verbose --no-cpu-baseline --no-ref-gpu --no-latency --steps 5 \
   > gpurun_out/r2c_extra2.json 2> gpurun_out/r2c_extra2.err; echo "exit $?"; grep -v "^\[tune\]" gpurun_out/r2c_extra2.err | tail -n 8
echo "== extra=2 cfg3 verbose overlapped"
PN_CONV_TUNE_EXTRA=2 PN_CONV_TUNING_FILE=none timeout 100 python bench.py --workload cfg3 --mode overlapped --verbose --no-cpu-baseline --no-ref-gpu --no-latency --steps 5 \
   > gpurun_out/r2c_extra2o.json 2> gpurun_out/r2c_extra2o.err; echo "exit $?"; grep -v "^\[tune\]" gpurun_out/r2c_extra2o.err | tail -n 8
echo "== default bench"
PN_TUNING_DUMP=gpurun_out/tuning_default.txt timeout 600 python bench.py --verbose > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err; echo "exit $?"
tail -n 5 gpurun_out/r2c_bench.err; cat gpurun_out/r2c_bench.json
echo "== cfg3 single GPU"
PN_TUNING_DUMP=gpurun_out/tuning_cfg3.txt timeout 300 python bench.py --workload cfg3 --no-cpu-baseline --no-ref-gpu --no-latency > gpurun_out/r2c_cfg3.json 2> gpurun_out/r2c_cfg3.err; echo "exit $?"
cut -c1-500 gpurun_out/r2c_cfg3.json
