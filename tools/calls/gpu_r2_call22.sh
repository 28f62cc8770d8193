#!/bin/bash
# Round 2, call 22: weights-resident conv mode: parity (bit-identical to the streaming launch), then timings on the layers whose
# weight slab fits beside the ring.  force word: bn | 0x4000 (no pair) | 1 << 16 (no split) | opt << 24 (1 two CTAs, 4 resident).
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_conv_gpu.py -m gpu -x -q -k "resident" 2>&1 | tail -n 5
out=gpurun_out/r2t_resident.txt
: > $out
run() { echo -n "force=$(printf 0x%x $4)  " >> $out; timeout 60 python tools/conv_one.py $1 $2 $3 20 $4 >> $out 2>&1 || echo "$1 force $4: FAILED/timeout" >> $out; }
for shape in A.res2.conv1 A.res2.conv2 A.res2.conv3 A.res3.conv3 A.res4.conv3 A.fpn_lat2; do
  for bn in 64 128; do
    for opt in 0 1 4 5; do
      run $shape bf16 32 $(( bn | 0x4000 | (1 << 16) | (opt << 24) ))
    done
  done
done
grep -v "^Traceback\|^  File" $out | cut -c1-150
