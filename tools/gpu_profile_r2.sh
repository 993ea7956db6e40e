# round-2 evidence: launch list of the bench command (ncu, gpu__time_duration), full captures of the two attention kernels
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_x3.csv python bench.py --steps 2 --warmup 1 --no-extras --no-cpu-baseline > gpurun_out/launch_bench.log 2>&1
tail -2 gpurun_out/launch_bench.log
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_attn_bwd_p -c 1 -f -o gpurun_out/r02_attn_bwd_p python tools/prof_attn.py 4 1024 200 1 > gpurun_out/ncu_bwd_p.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_attn_fwd_p -c 1 -f -o gpurun_out/r02_attn_fwd_p python tools/prof_attn.py 4 1024 200 1 > gpurun_out/ncu_fwd_p.log 2>&1
tail -2 gpurun_out/ncu_fwd_p.log
