mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_x3.csv python bench.py --steps 2 --warmup 1 --no-extras --no-cpu-baseline > gpurun_out/launch_bench.log 2>&1
tail -c 300 gpurun_out/launch_bench.log
