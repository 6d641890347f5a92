for mb in 4 6 8; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -fmad=false -Xcompiler -fPIC -shared -DCS_MAP_MIN_CTAS=$mb -o cooperative-search_b200/csrc/libcoopsearch.so cooperative-search_b200/csrc/runtime.cu cooperative-search_b200/csrc/flight.cu cooperative-search_b200/csrc/search.cu 2>/dev/null
  touch cooperative-search_b200/csrc/libcoopsearch.so
  for lpe in 0 32; do
  CS_BENCH_LPE=$lpe python bench.py --workload c4 --steps 100 --warmup 5 --no-extra 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('minb=$mb c4 lpe=$lpe us/launch %.2f value %.3e frac %.4f' % (d['roofline']['us_per_launch'], d['value'], d['roofline']['frac']))"
  done
done
python bench.py --workload c5 --steps 30 --warmup 3 --no-extra 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('c5 us/launch %.1f value %.3e frac %.3f e2e %.3e' % (d['roofline']['us_per_launch'], d['value'], d['roofline']['frac'], d['e2e']['value']))"
python -m pytest tests/test_gpu_search.py -q -m gpu 2>&1 | tail -2
