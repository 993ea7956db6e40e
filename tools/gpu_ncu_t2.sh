mkdir -p gpurun_out
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_attn_bwd_t2 -c 1 -f -o gpurun_out/r02_attn_bwd_t2 python tools/prof_attn.py 5 1024 200 1 > gpurun_out/ncu_bwd_t2.log 2>&1
tail -2 gpurun_out/ncu_bwd_t2.log
