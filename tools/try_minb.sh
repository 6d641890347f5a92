for mb in 7 8; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -fmad=false -Xcompiler -fPIC -shared -DCS_MAP_MIN_CTAS=$mb -o cooperative-search_b200/csrc/libcoopsearch.so cooperative-search_b200/csrc/runtime.cu cooperative-search_b200/csrc/flight.cu cooperative-search_b200/csrc/search.cu 2>/dev/null
  touch cooperative-search_b200/csrc/libcoopsearch.so
  python bench.py --workload c4 --steps 200 --warmup 5 --no-extra 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('minb=$mb c4 us/launch %.2f value %.3e frac %.4f' % (d['roofline']['us_per_launch'], d['value'], d['roofline']['frac']))"
done
