#!/bin/bash
# Round 2, multi-GPU call (run with gpurun --gpus N): peer-memory gather test, scaling bench lines with the gather in the step.
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r02_topo_n$N.txt 2>&1
timeout 600 python -m pytest tests/test_gather_gpu.py -x -q 2>&1 | tail -n 5
run() { tag=$1; shift
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N "$@" \
     > gpurun_out/r02_bench_n${N}_$tag.json 2> gpurun_out/r02_bench_n${N}_$tag.err; echo "n=$N $tag exit $?"
  tail -n 3 gpurun_out/r02_bench_n${N}_$tag.err | cut -c1-300; cut -c1-900 gpurun_out/r02_bench_n${N}_$tag.json
}
run peer --steps 20
run nccl --steps 20 --gather nccl
run nogather --steps 20 --gather none
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus $N --impl reference --steps 2 --warmup 0 | cut -c1-300
