#!/bin/bash
# Round 2, final tree on two GPUs (gpurun --gpus 2): peer-memory gather test and the default bench line exactly as the driver
# launches it (32 frames per GPU, gather inside the step, nested configs[3] block).
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gather_gpu.py -m gpu -x -q 2>&1 | tail -n 3
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 3 \
   > gpurun_out/r02_bench_n2_final_tree.json 2> gpurun_out/r02_bench_n2_final_tree.err; echo "n=2 exit $?"
tail -n 2 gpurun_out/r02_bench_n2_final_tree.err | cut -c1-300; cut -c1-900 gpurun_out/r02_bench_n2_final_tree.json
