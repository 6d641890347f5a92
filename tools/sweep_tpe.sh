#!/bin/bash
# thread-per-env kernel: CTA size sweep on the GPU box (variant builds), c3 / c2w / c2
mkdir -p gpurun_out /tmp/csv
cd cooperative-search_b200/csrc
for v in "32 16" "128 4" "256 2"; do
  set -- $v
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -fmad=false -Xcompiler -fPIC -shared -DCS_TPE_THREADS=$1 -DCS_TPE_MIN_CTAS=$2 -o /tmp/csv/lib_$1.so runtime.cu flight.cu search.cu 2>/dev/null &
done
wait
cd ../..
{
for v in "32 16" "128 4" "256 2"; do
  set -- $v
  echo "== threads per CTA $1 (min ctas $2)"
  COOPSEARCH_LIB=/tmp/csv/lib_$1.so python tools/sweep_step.py c3 c2w c2 2>&1 | grep value
done
} > gpurun_out/sweep_tpe.log 2>&1
cat gpurun_out/sweep_tpe.log
