timeout 600 python -m pytest tests/test_gpu_attn.py -x -q -k "tcgen05_t2 or tcgen05_p" 2>&1 | tail -5
timeout 120 python tools/prof_attn.py 5 1024 200 5 2>&1 | tail -1
timeout 120 python tools/prof_attn.py 4 1024 200 5 2>&1 | tail -2
