#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_policy.py tests/test_gpu_replay.py tests/test_gpu_flight_map.py tests/test_gpu_rollout.py tests/test_gpu_known_answers.py -x -q -m gpu > gpurun_out/e_tests.txt 2>&1
tail -25 gpurun_out/e_tests.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/e_smoke.txt 2>&1
tail -5 gpurun_out/e_smoke.txt
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/e_bench_c2.json 2> gpurun_out/e_bench_c2.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/e_bench_c2.json").read().strip().splitlines()[-1])
print("value", d["value"], "frac", d["roofline"]["frac"])
print("e2e", json.dumps(d["e2e"])[:600])
print("policy", json.dumps(d["extra"].get("policy"))[:2500])
print("c4", json.dumps(d["extra"].get("c4"))[:900])
PY
tail -5 gpurun_out/e_bench_c2.err
