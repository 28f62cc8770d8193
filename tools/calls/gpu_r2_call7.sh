#!/bin/bash
# Round 2, call 7: what is the stall made of?  12 runs per variant with the forward-ordering event disabled; ROIAlign timing.
mkdir -p gpurun_out
B="python bench.py --workload cfg3 --mode overlapped --no-cpu-baseline --no-ref-gpu --no-latency --no-profile --steps 5"
variant() { name=$1; shift; fails=0
  for i in 1 2 3 4 5 6 7 8 9 10 11 12; do
    env PN_DEBUG_NO_FWD_ORDER=1 PN_CONV_TUNE_EXTRA=2 PN_CONV_TUNING_FILE=none "$@" timeout 30 $B > gpurun_out/p7_$name.json 2> gpurun_out/p7_$name.err || fails=$((fails+1))
  done
  echo "variant $name: $fails of 12 runs did not finish"
}
variant base
variant pdl_only_in_graph PN_DEBUG_PDL_ONLY_IN_GRAPH=1
variant tmem_after_wait PN_DEBUG_TMEM_AFTER_WAIT=1
variant no_lanes PN_DEBUG_NO_LANES=1
timeout 300 python tools/mrcnn_profile.py 8 bf16 > gpurun_out/p7_ops_mrcnn_b8_bf16.txt 2>&1; grep -E "roi_align|^#" gpurun_out/p7_ops_mrcnn_b8_bf16.txt
timeout 600 python -m pytest tests/test_maskrcnn_gpu.py tests/test_prednet_gpu.py -x -q 2>&1 | tail -n 5
