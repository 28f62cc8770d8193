#!/bin/bash
# Round 2, call 20: experiment - CTA pairs combined with two CTAs per SM (BN = 128) on the L2-bound 1x1 layers, against the
# table's choice.  force word: bn | 0x2000 (pair) | 1 << 24 (two CTAs per SM).
mkdir -p gpurun_out
F=$((128 | 0x2000 | (1 << 24)))
out=gpurun_out/r2p_pair_opt1.txt
: > $out
for shape in A.res4.conv3 A.res3.conv3 A.res2.conv3 A.res5.conv3 A.res4.conv1 A.res3.conv1 A.res5.conv1 A.fpn_lat2; do
  for f in 0 $F $((128 | (1 << 24))) $((128 | 0x2000)) $((256 | 0x2000)); do
    echo -n "force=$f  " >> $out
    timeout 60 python tools/conv_one.py $shape bf16 32 20 $f >> $out 2>&1 || echo "$shape force $f: FAILED/timeout" >> $out
  done
done
cat $out | grep -v "^Traceback\|^  File" | cut -c1-160
