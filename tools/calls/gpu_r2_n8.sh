#!/bin/bash
# Round 2, 8-GPU call (gpurun --gpus 8): the default bench line at N = 8 exactly as the driver launches it, then N = 4 on the same box.
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r02_topo_n8.txt 2>&1
for N in 8 4; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 20 --warmup 3 \
     > gpurun_out/r02_bench_n${N}_peer.json 2> gpurun_out/r02_bench_n${N}_peer.err; echo "n=$N exit $?"
  tail -n 2 gpurun_out/r02_bench_n${N}_peer.err | cut -c1-300
  grep "^{" gpurun_out/r02_bench_n${N}_peer.json | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('N=%d value %.1f fps %.2f ms e2e %.1f gather_ms %s cfg3 %.1f fps (%.2f ms, gather %s) clocks %s' % (d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d['run']['gather_ms'], d['cfg3']['value'], d['cfg3']['ms_per_step'], d['cfg3']['gather_ms'], d['clocks']))"
done
