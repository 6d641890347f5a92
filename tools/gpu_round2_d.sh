#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_flight_easy.py tests/test_gpu_policy.py tests/test_gpu_flight_map.py -x -q -m gpu > gpurun_out/d_tests.txt 2>&1
tail -25 gpurun_out/d_tests.txt
timeout 300 python bench.py --workload c4 --no-extra --steps 100 --warmup 5 > gpurun_out/d_bench_c4.json 2> gpurun_out/d_bench_c4.err
head -c 2500 gpurun_out/d_bench_c4.json; tail -5 gpurun_out/d_bench_c4.err
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/d_bench_c2.json 2> gpurun_out/d_bench_c2.err
head -c 6000 gpurun_out/d_bench_c2.json; tail -5 gpurun_out/d_bench_c2.err
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/d_bench_ref.json 2> gpurun_out/d_bench_ref.err
cat gpurun_out/d_bench_ref.json; tail -5 gpurun_out/d_bench_ref.err
