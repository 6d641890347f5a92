#!/bin/bash
mkdir -p gpurun_out
{
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python tools/sweep_step.py c2 c2g 2>&1 | grep value
CS_BENCH_LPE=4 python tools/sweep_step.py c2g 2>&1 | grep value
} > gpurun_out/sweep_step.log 2>&1
cat gpurun_out/sweep_step.log
