timeout 300 python -m pytest tests/test_gpu_attn.py -x -q -k "tcgen05_p and backward" 2>&1 | tail -15
timeout 120 python tools/prof_attn.py 4 1024 200 5 2>&1 | tail -3
timeout 120 python tools/prof_attn.py 3 1024 200 5 2>&1 | tail -3
