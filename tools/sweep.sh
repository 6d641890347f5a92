#!/bin/bash
for st in 1 4 8 16; do
  CS_BENCH_STREAMS=$st python bench.py --workload c2 --steps 300 --warmup 5 --no-extra 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('c2 streams=$st us/launch %.2f value %.3e frac %.4f e2e %.3e' % (d['roofline']['us_per_launch'], d['value'], d['roofline']['frac'], d['e2e']['value']))"
done
