#!/bin/bash
# Round 2: DRAM traffic of the N4 kernels against their algorithmic bytes (ncu, no clock control).
mkdir -p gpurun_out
timeout 240 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:k_map_sample --launch-skip 3 -c 4 --csv \
   --log-file gpurun_out/r02_map_dataset_ncu.csv python tools/map_dataset_profile.py > gpurun_out/r02_map_dataset_ncu.log 2>&1; echo "ncu exit $?"
grep -v "^==" gpurun_out/r02_map_dataset_ncu.csv | python -c "
import csv, sys
rows = list(csv.DictReader(sys.stdin))
from collections import OrderedDict
agg = OrderedDict()
for r in rows:
    k = (r['ID'], r['Kernel Name'].split('(')[0], r['Grid Size'])
    agg.setdefault(k, {})[r['Metric Name']] = (float(r['Metric Value'].replace(',', '')), r['Metric Unit'])
for k, m in agg.items():
    print(k[1], 'grid', k[2], ' '.join(f'{n}={v[0]:.1f}{v[1]}' for n, v in m.items()))
"
