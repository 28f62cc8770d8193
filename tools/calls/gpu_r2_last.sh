#!/bin/bash
# Round 2, last call: smoke and both bench arms (one stdout line each) on the final commit.
mkdir -p gpurun_out
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 4
timeout 200 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/last_ref.out 2>/dev/null; echo "ref exit $? lines $(wc -l < gpurun_out/last_ref.out)"; cut -c1-260 gpurun_out/last_ref.out
timeout 300 python bench.py --steps 10 --warmup 3 --no-ref-gpu --no-latency --no-scaling-base > gpurun_out/last_ours.out 2>/dev/null; echo "ours exit $? lines $(wc -l < gpurun_out/last_ours.out)"
python -c "
import json; d=json.loads(open('gpurun_out/last_ours.out').read()); print(d['value'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['traffic_source'][:60], d['cpu_baseline']['value'], d['clocks'])"
