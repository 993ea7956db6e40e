mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus 2 --steps 20 --warmup 5 --ids realistic --no-extras --no-cpu-baseline > gpurun_out/bench_n2_real.json 2> gpurun_out/bench_n2_real.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_n2_real.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d.get('e2e'))
print(json.dumps(d.get('dp_check'))[:400])
print(json.dumps(d.get('sharded_20M'))[:300])
PY
tail -3 gpurun_out/bench_n2_real.err
