#!/bin/bash
# Round 2, call 9: conv parity incl. the direct-residual epilogue, soak of the tensor-memory fix, single-layer timings
# (res4.conv3 / res3.conv3 / res2.conv3 at batch 32 bf16 across launch options), full GPU tests, benches + tuning tables.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_conv_gpu.py -x -q 2>&1 | tail -n 6
run() { python tools/conv_one.py $1 $2 $3 20 $4 2>&1 | tail -1 | sed "s/\$/   force=$4/"; }
for shape in A.res4.conv3 A.res3.conv3 A.res2.conv3 A.res5.conv3; do
  for bn in 128 256; do
    for opt in 0 1 4 5; do
      if [ $bn = 256 ] && [ $opt = 1 -o $opt = 5 ]; then continue; fi
      f=$(printf "0x%x" $(( bn | 0x4000 | (1<<16) | (opt<<24) )))
      run $shape bf16 32 $f
    done
  done
  for opt in 0 4; do f=$(printf "0x%x" $(( 256 | 0x2000 | (opt<<24) ))); run $shape bf16 32 $f; done
done > gpurun_out/r2f_res_direct_b32.txt 2>&1
cat gpurun_out/r2f_res_direct_b32.txt
B="python bench.py --workload cfg3 --mode overlapped --no-cpu-baseline --no-ref-gpu --no-latency --no-profile --steps 5"
fails=0
for i in $(seq 1 20); do
  env PN_DEBUG_NO_FWD_ORDER=1 PN_CONV_TUNING_FILE=none timeout 40 $B > gpurun_out/p9_soak.json 2> gpurun_out/p9_soak.err || fails=$((fails+1))
done
echo "soak (alloc after wait, no event, all extra tuner modes on): $fails of 20 runs did not finish" | tee -a gpurun_out/r2f_soak.txt
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2f_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2f_pytest.log
tail -n 25 gpurun_out/r2f_pytest.log
PN_CONV_TUNING_FILE=none PN_TUNING_DUMP=gpurun_out/tuning_cfg2_cfg1.txt timeout 600 python bench.py --verbose > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err; echo "bench exit $?"
grep -v "^\[tune\]" gpurun_out/r2f_bench.err | tail -n 4; cat gpurun_out/r2f_bench.json
PN_CONV_TUNING_FILE=none PN_TUNING_DUMP=gpurun_out/tuning_cfg3.txt timeout 300 python bench.py --workload cfg3 --no-cpu-baseline --no-ref-gpu --no-latency > gpurun_out/r2f_cfg3.json 2> gpurun_out/r2f_cfg3.err; echo "cfg3 exit $?"
cut -c1-600 gpurun_out/r2f_cfg3.json
timeout 300 python tools/mrcnn_profile.py 32 bf16 > gpurun_out/r2f_ops_mrcnn_b32_bf16.txt 2>&1; tail -n 2 gpurun_out/r2f_ops_mrcnn_b32_bf16.txt
