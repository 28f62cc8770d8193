#!/bin/bash
# CTA-0 timelines of single-wave conv layers.  The library must be built with the timeline probes:
#   make -C peanut_b200/csrc clean && make -C peanut_b200/csrc -j8 NVCCFLAGS_EXTRA=-DPN_CONV_TIMELINE
# (rebuild without the flag afterwards).  Output: clock64 stamps of CTA 0 per tile (producer, MMA issuer, epilogue).
for spec in "A.res4.conv3 tf32 1 8 64" "A.res4.conv3 tf32 1 8 256" "A.res4.conv3 bf16 8 8 0" "A.res4.conv1 tf32 1 8 0"; do
  echo "== $spec"; PN_CONV_DBG=1 python tools/conv_one.py $spec 2>&1 | tail -12
done
