#!/bin/bash
# Epilogue attribution by elimination (library built with -DPN_CONV_TIMELINE): PN_EPI_SKIP bits 1 TMA store, 2 st.shared,
# 4 residual, 8 bias loads, 16 no tcgen05.ld double buffering.  Prints the launch time and the timeline of CTA 0's 4th tile.
for spec in "A.res4.conv3 bf16 8 8 128" "A.res4.conv1 bf16 8 8 0x4100" "A.res4.conv3 tf32 1 8 128"; do
  for skip in 0 1 2 3 4 8 15 16; do
    echo "== $spec skip=$skip"; PN_EPI_SKIP=$skip PN_CONV_DBG=1 python tools/conv_one.py $spec 2>&1 | grep "^tile  [13]\|TF/s"
  done
done
