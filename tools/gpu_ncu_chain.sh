mkdir -p gpurun_out
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"k_(qkv_bwd|ffn_bwd|proj_ffn|ln_qkv)_x3" --launch-skip 0 -c 16 -f -o gpurun_out/r02_chain_x3_final python bench.py --steps 1 --warmup 1 --no-extras --no-cpu-baseline > gpurun_out/ncu_chain.log 2>&1
tail -2 gpurun_out/ncu_chain.log | cut -c1-200
