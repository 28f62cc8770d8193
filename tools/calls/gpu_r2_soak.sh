#!/bin/bash
# Round 2: soak of the three bench workloads at HEAD (the round-1 "cfg3 stall" was a 2-in-4 event): 12 fresh processes each of
# cfg3 and cfg2 with a 120 s watchdog; prints value per run and counts the runs that did not finish.
mkdir -p gpurun_out
out=gpurun_out/r02_soak.txt
: > $out
F="--no-cpu-baseline --no-ref-gpu --no-latency --no-scaling-base --no-profile --steps 10 --warmup 3"
for wl in cfg3 cfg2; do
  for i in $(seq 1 12); do
    timeout 120 python bench.py $F --workload $wl > gpurun_out/soak_line.json 2>/dev/null; rc=$?
    v=$(python -c "import json;d=json.load(open('gpurun_out/soak_line.json'));print('%.1f fps %.2f ms' % (d['value'], d['ms_per_step']))" 2>/dev/null)
    echo "$wl run $i rc=$rc $v" >> $out
  done
done
cat $out | awk '{print $1, $3}' | sort | uniq -c
