mkdir -p gpurun_out
T=r2s3mg8
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"
tail -n 3 gpurun_out/${T}_bench.err
python - <<PY
import json
d = json.loads(open('gpurun_out/r2s3mg8_bench.json').read().strip().splitlines()[-1])
print('N', d['n_gpus'], 'ms_per_step', d['ms_per_step'], 'value', d['value'], 'e2e ms', d['e2e']['ms_per_step'])
print(json.dumps(d.get('sharded_batch'), indent=1))
PY
