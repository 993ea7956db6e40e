mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_dp.py -x -q -m gpu 2>&1 | tail -3
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_n2.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d.get('e2e'))
print(json.dumps(d.get('dp_check'))[:1200])
print(json.dumps(d.get('sharded_20M'))[:600])
PY
