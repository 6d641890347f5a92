#!/bin/bash
# first GPU pass of round 2: map tests first (fast signal), then the whole GPU suite, then c4 / c2 timings
mkdir -p gpurun_out
python -m pytest tests/test_gpu_flight_map.py -x -q -m gpu 2>&1 | tail -25 > gpurun_out/a_map_tests.txt
cat gpurun_out/a_map_tests.txt
python -m pytest tests -x -q -m gpu 2>&1 | tail -25 > gpurun_out/a_all_tests.txt
cat gpurun_out/a_all_tests.txt
python bench.py --workload c4 --no-extra --steps 400 --warmup 20 > gpurun_out/a_bench_c4.json 2> gpurun_out/a_bench_c4.err
cat gpurun_out/a_bench_c4.json | head -c 3000; tail -5 gpurun_out/a_bench_c4.err
python bench.py --no-extra --steps 400 --warmup 20 > gpurun_out/a_bench_c2.json 2> gpurun_out/a_bench_c2.err
cat gpurun_out/a_bench_c2.json | head -c 1500
