#!/bin/bash
mkdir -p gpurun_out
{
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python tools/sweep_step.py c2 c3 c2w c4 2>&1 | grep value
} > gpurun_out/sweep_step.log 2>&1
cat gpurun_out/sweep_step.log
