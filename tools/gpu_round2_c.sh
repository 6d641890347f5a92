#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_flight_map.py tests/test_gpu_edge_cases.py tests/test_gpu_search.py -x -q -m gpu 2>&1 | tail -8 > gpurun_out/c_tests.txt
cat gpurun_out/c_tests.txt
{
for v in l16m12u2 l8m7u4 l8m10u2 l8m8u4 l16m16u1 l16m14u1; do
  COOPSEARCH_LIB=cooperative-search_b200/csrc/variants/$v.so python tools/exp_c4.py 0 0
  COOPSEARCH_LIB=cooperative-search_b200/csrc/variants/$v.so python tools/exp_c4.py 0 1
done
} 2>&1 | grep -v "^Init" > gpurun_out/c_exp.txt
cat gpurun_out/c_exp.txt
