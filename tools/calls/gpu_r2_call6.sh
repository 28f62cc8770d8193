#!/bin/bash
# Round 2, call 6: is the forward-ordering event what removed the stall?  8 runs without it, then GPU tests of the new code.
mkdir -p gpurun_out
for i in 1 2 3 4 5 6 7 8; do
  PN_DEBUG_NO_FWD_ORDER=1 PN_CONV_TUNE_EXTRA=2 PN_CONV_TUNING_FILE=none timeout 40 python bench.py --workload cfg3 --mode overlapped --no-cpu-baseline --no-ref-gpu --no-latency --no-profile --steps 5 > gpurun_out/p6_$i.json 2> gpurun_out/p6_$i.err; echo "no-fwd-order run $i exit $?"
done
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2d_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2d_pytest.log
tail -n 15 gpurun_out/r2d_pytest.log
