#!/bin/bash
# lanes-per-env sweep of the flight workloads (tuning aid; prints us per launch)
for wl in c2 c3 c4; do
  for lpe in 0 32; do
    CS_BENCH_LPE=$lpe python bench.py --workload $wl --steps 100 --warmup 5 --no-extra 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$wl lpe=$lpe(%s) us/launch %.2f value %.3e frac %.4f e2e %.3e' % (d['config']['lanes_per_env'], d['roofline']['us_per_launch'], d['value'], d['roofline']['frac'], d['e2e']['value']))"
  done
done
