#!/bin/bash
# Round 2, call 4: force two CTAs per SM on every eligible layer and look for the stall; end-to-end parity probe.
mkdir -p gpurun_out
run() { name=$1; shift
  echo "== $name"
  ( timeout 90 env "$@" > gpurun_out/p4_$name.log 2>&1; echo "exit $?" >> gpurun_out/p4_$name.log )
  grep -v "^\[tune\]\|force-opt1" gpurun_out/p4_$name.log | tail -n 4 | cut -c1-300
  echo "forced layers: $(grep -c force-opt1 gpurun_out/p4_$name.log)"
}
B="python bench.py --workload cfg3 --verbose --no-cpu-baseline --no-ref-gpu --no-latency --steps 5"
run all_dep PN_CONV_FORCE_OPT1=all PN_CONV_TUNE_LOG=1 $B
run all_ovl PN_CONV_FORCE_OPT1=all PN_CONV_TUNE_LOG=1 $B --mode overlapped
run all_dep_sync PN_CONV_FORCE_OPT1=all PN_DEBUG_SYNC_EACH=1 $B --no-profile
run tune2_dep PN_CONV_TUNE_EXTRA=2 PN_CONV_TUNING_FILE=none PN_CONV_TUNE_LOG=1 $B
run tune2_ovl PN_CONV_TUNE_EXTRA=2 PN_CONV_TUNING_FILE=none PN_CONV_TUNE_LOG=1 $B --mode overlapped
run tune2_ovl_b PN_CONV_TUNE_EXTRA=2 PN_CONV_TUNING_FILE=none PN_CONV_TUNE_LOG=1 $B --mode overlapped
run all_cfg2 PN_CONV_FORCE_OPT1=all python bench.py --workload cfg2 --verbose --no-cpu-baseline --no-ref-gpu --no-latency --steps 5
echo "== parity probe"
timeout 600 python tools/e2e_parity_probe.py 0.3 > gpurun_out/p4_parity.log 2>&1; tail -n 12 gpurun_out/p4_parity.log
