#!/bin/bash
# Round 2, verification of the final tree on one GPU: full GPU tests, smoke, whole-chain margins in the three modes.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_final3_pytest.log 2>&1; echo "pytest exit $?"; tail -n 3 gpurun_out/r02_final3_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 5
timeout 300 python tools/whole_chain_probe.py > gpurun_out/r02_whole_chain_modes.txt 2>&1; echo "probe exit $?"; cat gpurun_out/r02_whole_chain_modes.txt
