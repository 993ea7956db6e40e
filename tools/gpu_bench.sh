mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_tc.py -x -q -m gpu 2>&1 | tail -4
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_n1.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d.get('e2e'))
for e in d.get('kernel_breakdown',[])[:10]: print(e['kernel'], round(e['ms_per_step'],3), round(e.get('frac',0),3))
PY
