#!/bin/bash
for spec in "A.res4.conv3 tf32 1 8 256" "A.res4.conv3 bf16 8 8 0" "A.res4.conv1 bf16 8 8 0"; do
  echo "== $spec"; PN_CONV_DBG=1 python tools/conv_one.py $spec 2>&1 | grep -v "^launch" | tail -5
done
