#!/bin/bash
# Round 2: strict-parity (fp32 = 3 x tf32) mode: probe of the margins behind the new tests' tolerances.
mkdir -p gpurun_out
timeout 600 python tools/fp32_parity_probe.py conv prednet maskrcnn > gpurun_out/r02_fp32_probe.txt 2>&1; echo "probe exit $?"
tail -n 60 gpurun_out/r02_fp32_probe.txt
