#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_flight_map.py tests/test_gpu_flight_easy.py tests/test_gpu_edge_cases.py tests/test_gpu_parity_large.py -x -q -m gpu 2>&1 | tail -15 > gpurun_out/b_tests.txt
cat gpurun_out/b_tests.txt
{
python tools/exp_c4.py 0 0
python tools/exp_c4.py 0 1
python tools/exp_c4.py 8 0
for v in m14u2 m12u2 m10u4 m14u4; do
  COOPSEARCH_LIB=cooperative-search_b200/csrc/variants/$v.so python tools/exp_c4.py 0 0
  COOPSEARCH_LIB=cooperative-search_b200/csrc/variants/$v.so python tools/exp_c4.py 0 1
done
} 2>&1 | grep -v "^Init" > gpurun_out/b_exp.txt
cat gpurun_out/b_exp.txt
