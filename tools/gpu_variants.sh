mkdir -p gpurun_out
timeout 600 python bench.py --steps 20 --warmup 5 --dr --no-extras --no-cpu-baseline > gpurun_out/bench_n1_dr.json 2> gpurun_out/bench_n1_dr.err
timeout 600 python bench.py --steps 20 --warmup 5 --ids realistic --no-extras --no-cpu-baseline > gpurun_out/bench_n1_real.json 2> gpurun_out/bench_n1_real.err
python - <<'PY'
import json
for f in ('bench_n1_dr','bench_n1_real'):
    d=json.loads(open(f'gpurun_out/{f}.json').read().strip().splitlines()[-1])
    print(f, d['value'], d['ms_per_step'], d['e2e']['value'])
PY
