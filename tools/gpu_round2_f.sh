#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_flight_easy.py tests/test_gpu_policy.py -x -q -m gpu > gpurun_out/f_tests.txt 2>&1
tail -12 gpurun_out/f_tests.txt
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/f_bench_c2.json 2> gpurun_out/f_bench_c2.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/f_bench_c2.json").read().strip().splitlines()[-1])
print("value", d["value"], "frac", d["roofline"]["frac"], "us", d["roofline"]["us_per_launch"])
print("e2e", json.dumps(d["e2e"])[:400])
print("policy", json.dumps(d["extra"].get("policy"))[:2500])
PY
tail -5 gpurun_out/f_bench_c2.err
