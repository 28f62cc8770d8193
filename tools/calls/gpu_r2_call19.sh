#!/bin/bash
# Round 2, call 19: fuse kernel with register accumulators - mapper / pipeline tests, mapper timings and launch lists.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_mapper_gpu.py tests/test_pipeline_gpu.py tests/test_agent_prediction_gpu.py -m gpu -x -q > gpurun_out/r2o_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2o_pytest.log
tail -n 6 gpurun_out/r2o_pytest.log
for E in 1 8 32; do timeout 200 python tools/mapper_profile.py $E > gpurun_out/r2o_mapper_e$E.txt 2>&1; cat gpurun_out/r2o_mapper_e$E.txt | tail -n 2; done
for E in 1 8 32; do
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2o_mapper_launches_e$E.csv python tools/mapper_profile.py $E > /dev/null 2>&1
  grep k_fuse gpurun_out/r2o_mapper_launches_e$E.csv | tail -n 1 | cut -d, -f5,12- 
done
