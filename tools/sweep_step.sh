#!/bin/bash
mkdir -p gpurun_out
{
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
CS_TPE_K=1 python -m pytest tests/test_gpu_flight_easy.py tests/test_gpu_flight_map.py tests/test_gpu_rollout.py -m gpu -x -q 2>&1 | tail -3
python tools/sweep_step.py c2 c3 c2w c4 2>&1 | grep value
} > gpurun_out/sweep_step.log 2>&1
cat gpurun_out/sweep_step.log
