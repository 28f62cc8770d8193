#!/bin/bash
# One GPU call: parity tests, bench lines, per-op profiles. Outputs under gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench_cfg1.json 2> gpurun_out/bench_cfg1.err
timeout 600 python bench.py --workload cfg3 --no-cpu-baseline > gpurun_out/bench_cfg3.json 2> gpurun_out/bench_cfg3.err
timeout 600 python bench.py --workload cfg2 --no-cpu-baseline --steps 10 > gpurun_out/bench_cfg2.json 2> gpurun_out/bench_cfg2.err
timeout 300 python tools/mrcnn_profile.py 1 tf32 > gpurun_out/ops_mrcnn_b1_tf32.txt 2>&1
timeout 300 python tools/mrcnn_profile.py 8 bf16 > gpurun_out/ops_mrcnn_b8_bf16.txt 2>&1
timeout 300 python tools/prednet_profile.py 1 24 240 tf32 > gpurun_out/ops_pred_b1_tf32.txt 2>&1
timeout 300 python tools/prednet_profile.py 8 24 240 bf16 > gpurun_out/ops_pred_b8_bf16.txt 2>&1
tail -3 gpurun_out/pytest_gpu.log; cat gpurun_out/bench_cfg1.json gpurun_out/bench_cfg3.json gpurun_out/bench_cfg2.json
