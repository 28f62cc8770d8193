#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 300 python tools/mrcnn_profile.py 1 tf32 > gpurun_out/ops_mrcnn_b1_tf32.txt 2>&1
timeout 300 python tools/prednet_profile.py 1 24 240 tf32 > gpurun_out/ops_pred_b1_tf32.txt 2>&1
timeout 300 python tools/mrcnn_profile.py 8 bf16 > gpurun_out/ops_mrcnn_b8_bf16.txt 2>&1
grep -E "graph replay|eager sum" gpurun_out/ops_mrcnn_b1_tf32.txt gpurun_out/ops_pred_b1_tf32.txt gpurun_out/ops_mrcnn_b8_bf16.txt
for w in cfg1 cfg3; do python bench.py --no-cpu-baseline --no-profile --workload $w 2>/dev/null | python -c "import json,sys; j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$w', round(j['value'],1), 'fps e2e', round(j['e2e']['value'],1), j['clocks'])"; done
