#!/bin/bash
# Round 2, multi-GPU call (run with gpurun --gpus N): peer-memory gather test, the default bench line (32 frames per GPU with the
# gather in the step + the nested configs[3] block), the NCCL-gather and no-gather variants of the headline, the reference arm.
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r02_topo_n$N.txt 2>&1
timeout 600 python -m pytest tests/test_gather_gpu.py tests/test_mapper_gpu.py -x -q 2>&1 | tail -n 5
run() { tag=$1; shift
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N "$@" \
     > gpurun_out/r02_bench_n${N}_$tag.json 2> gpurun_out/r02_bench_n${N}_$tag.err; echo "n=$N $tag exit $?"
  tail -n 3 gpurun_out/r02_bench_n${N}_$tag.err | cut -c1-300; cut -c1-1200 gpurun_out/r02_bench_n${N}_$tag.json
}
run peer --steps 20
run nccl --steps 20 --gather nccl --no-scaling-base
run nogather --steps 20 --gather none --no-scaling-base
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus $N --impl reference --steps 2 --warmup 0 | cut -c1-300
