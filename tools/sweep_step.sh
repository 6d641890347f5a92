#!/bin/bash
mkdir -p gpurun_out
{
python tools/sweep_step.py c2 c2s 2>&1 | grep value
python bench.py --steps 400 --warmup 10 --no-extra 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('bench c2', d['value'], d['ms_per_step'], d['roofline']['frac'], d['gpu_launches'], d['e2e']['value'])"
} > gpurun_out/sweep_step.log 2>&1
cat gpurun_out/sweep_step.log
