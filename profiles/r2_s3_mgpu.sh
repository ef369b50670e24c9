# multi-GPU call: NCCL parity test of the sharded path + bench.py at N = $1 (torchrun, one rank per GPU)
N=${1:-2}
mkdir -p gpurun_out
T=r2s3mgb${N}
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -n 8
timeout 600 python -m pytest tests/test_distributed_gpu.py tests/test_gpu_parity.py -k "nccl or sharded" -q --timeout 500 -p no:cacheprovider > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?"
tail -n 4 gpurun_out/${T}_pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"
tail -n 3 gpurun_out/${T}_bench.err
python - <<PY
import json
d = json.loads(open('gpurun_out/${T}_bench.json').read().strip().splitlines()[-1])
print('N', d['n_gpus'], 'ms_per_step', d['ms_per_step'], 'value', d['value'], 'e2e ms', d['e2e']['ms_per_step'])
print(json.dumps(d.get('sharded_batch'), indent=1))
PY
