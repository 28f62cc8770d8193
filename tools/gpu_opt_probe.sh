#!/bin/bash
# Launch-shape options on the epilogue-bound layers: force word = bn | 0x4000 (no pair) | 1<<16 (no split) | opt<<24
# (opt 1 = two CTAs per SM, 2 = bias through the tensor core, 3 = both).
run() { python tools/conv_one.py $1 $2 $3 20 $4 2>&1 | tail -1 | sed "s/\$/   force=$4/"; }
for opt in 0 1 2 3; do
  for bn in 128 64; do
    f=$(printf "0x%x" $(( bn | 0x4000 | (1<<16) | (opt<<24) )))
    run A.res4.conv3 bf16 8 $f
  done
done
for shape in A.res4.conv1 A.res3.conv3 A.res2.conv3 A.res3.conv1 A.res2.conv2 A.res3.conv2; do
  for bn in 128 64; do
    for opt in 0 1 3; do
      f=$(printf "0x%x" $(( bn | 0x4000 | (1<<16) | (opt<<24) )))
      run $shape bf16 8 $f
    done
  done
done
for opt in 0 1 2 3; do
  f=$(printf "0x%x" $(( 128 | 0x4000 | (1<<16) | (opt<<24) )))
  run A.res4.conv3 tf32 8 $f
done
