#!/bin/bash
# map kernel variants (builds on the GPU box), timed on c4
mkdir -p gpurun_out /tmp/csv
cd cooperative-search_b200/csrc
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -fmad=false -Xcompiler -fPIC -shared -DCS_MAP_PREFETCH -o /tmp/csv/lib_pf.so runtime.cu flight.cu search.cu 2>/dev/null &
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -fmad=false -Xcompiler -fPIC -shared -DCS_MAP_PREFETCH -DCS_MAP_MIN_CTAS=6 -o /tmp/csv/lib_pf6.so runtime.cu flight.cu search.cu 2>/dev/null &
wait
cd ../..
{
echo "== default"; python tools/exp_c4.py 300
echo "== L1 prefetch of the box rows"; COOPSEARCH_LIB=/tmp/csv/lib_pf.so python tools/exp_c4.py 300
COOPSEARCH_LIB=/tmp/csv/lib_pf.so python -m pytest tests/test_gpu_flight_map.py -m gpu -x -q 2>&1 | tail -2
echo "== L1 prefetch, 6 CTAs/SM"; COOPSEARCH_LIB=/tmp/csv/lib_pf6.so python tools/exp_c4.py 300
} > gpurun_out/sweep_map.log 2>&1
cat gpurun_out/sweep_map.log
