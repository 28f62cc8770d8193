#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_conv_gpu.py -m gpu -x -q 2>&1 | tail -3
for spec in "A.res4.conv1 tf32 1" "A.res4.conv2 tf32 1" "A.res5.conv1 tf32 1" "A.res5.conv2 tf32 1" "C.psp.bottleneck tf32 1" "C.l4.conv2 tf32 1"; do python tools/conv_one.py $spec 20; done
timeout 300 python tools/mrcnn_profile.py 1 tf32 > gpurun_out/ops_mrcnn_b1_tf32.txt 2>&1
timeout 300 python tools/prednet_profile.py 1 24 240 tf32 > gpurun_out/ops_pred_b1_tf32.txt 2>&1
grep -E "graph replay" gpurun_out/ops_mrcnn_b1_tf32.txt gpurun_out/ops_pred_b1_tf32.txt
python bench.py --no-cpu-baseline --no-profile 2>/dev/null | python -c "import json,sys; j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('cfg1', round(j['value'],1), 'fps e2e', round(j['e2e']['value'],1))"
