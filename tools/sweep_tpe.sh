#!/bin/bash
# thread-per-env kernel: register budget sweep on the GPU box (variant builds), K=1 on c3 / c2w / c4
mkdir -p gpurun_out /tmp/csv
cd cooperative-search_b200/csrc
for c in 8 10; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -fmad=false -Xcompiler -fPIC -shared -DCS_TPE_MIN_CTAS=$c -o /tmp/csv/lib$c.so runtime.cu flight.cu search.cu 2>/dev/null &
done
wait
cd ../..
{
python -m pytest tests/test_gpu_flight_easy.py -m gpu -x -q 2>&1 | tail -3
for c in 8 10; do
  echo "== min ctas $c"
  COOPSEARCH_LIB=/tmp/csv/lib$c.so CS_TPE_K=1 python tools/sweep_step.py c3 c2w 2>&1 | grep value
  COOPSEARCH_LIB=/tmp/csv/lib$c.so CS_TPE_K=1 python tools/exp_c4.py 200 2>&1 | tail -1
done
} > gpurun_out/sweep_tpe.log 2>&1
cat gpurun_out/sweep_tpe.log
CS_TPE_K=1 ncu --set full --clock-control none --import-source on -k regex:flight_tpe_kernel -s 12 -c 1 -f -o gpurun_out/prof_c2w_tpe python tools/profile_run.py c2w 16 > /dev/null 2>&1
