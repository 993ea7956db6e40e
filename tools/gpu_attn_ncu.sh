mkdir -p gpurun_out
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_attn_bwd_p -c 1 -f -o gpurun_out/r2_attn_bwd_p python tools/prof_attn.py 4 256 200 1 > gpurun_out/ncu_bwd_p.log 2>&1
tail -3 gpurun_out/ncu_bwd_p.log
