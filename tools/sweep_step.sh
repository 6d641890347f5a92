#!/bin/bash
mkdir -p gpurun_out
{
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for st in 1 2 4 8 16; do CS_BENCH_STREAMS=$st python tools/sweep_step.py c2 2>&1 | grep value; done
python tools/sweep_step.py c3 c2w c4 2>&1 | grep value
} > gpurun_out/sweep_step.log 2>&1
cat gpurun_out/sweep_step.log
