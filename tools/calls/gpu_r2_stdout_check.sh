#!/bin/bash
# bench.py under torchrun: exactly one line on stdout (the JSON), library chatter on stderr.
mkdir -p gpurun_out
B="--no-cpu-baseline --no-ref-gpu --no-latency --no-scaling-base --no-profile"
NCCL_DEBUG=VERSION timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 5 --warmup 3 $B \
   > gpurun_out/stdout_check.out 2> gpurun_out/stdout_check.err; echo "exit $?"
echo "stdout lines: $(wc -l < gpurun_out/stdout_check.out)"; python -c "
import json; d=json.loads(open('gpurun_out/stdout_check.out').read()); print('parsed', d['value'], d['n_gpus'], d['run']['gather_ms'])"
grep -c "NCCL version" gpurun_out/stdout_check.err
timeout 200 python bench.py --steps 3 --warmup 3 $B > gpurun_out/stdout_check1.out 2>/dev/null; echo "n=1 exit $? lines $(wc -l < gpurun_out/stdout_check1.out)"; cut -c1-200 gpurun_out/stdout_check1.out
