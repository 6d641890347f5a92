ncu --set full --clock-control none --import-source on -k regex:flight_kernel -s 40 -c 1 -f -o gpurun_out/prof_c4 python tools/profile_run.py c4 50 > /dev/null 2>&1
