#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/whole_chain_probe.py > gpurun_out/r02_whole_chain_modes.txt 2>&1; echo "exit $?"; tail -n 12 gpurun_out/r02_whole_chain_modes.txt
