make -C oracle -B liboracle.so >/dev/null 2>&1
python bench.py > gpurun_out/bench_r01.json 2> gpurun_out/bench_r01.err
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_r01_ref.json 2>> gpurun_out/bench_r01.err
mkdir -p gpurun_out
for wl in c3 c4 c5; do
  ncu --metrics gpu__time_duration.sum --clock-control none -s 10 -c 40 --csv --log-file gpurun_out/launches_$wl.csv python tools/profile_run.py $wl 30 > /dev/null 2>&1
done
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:flight_tpe_group_kernel -s 4 -c 8 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 2 --warmup 3 --no-extra > /dev/null 2>&1
tail -c 300 gpurun_out/bench_r01.err
wc -c gpurun_out/bench_r01.json gpurun_out/launches_*.csv
