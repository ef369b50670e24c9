# 2-GPU call: single-GPU test of the chunked expansion kernel, NCCL parity test of the sharded path, chunk sweep at config 5
N=${1:-2}
mkdir -p gpurun_out
T=r2s4mg${N}
timeout 400 python -m pytest tests/test_distributed_gpu.py tests/test_gpu_parity.py -k "nccl or sharded" -q --timeout 300 -p no:cacheprovider > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?"
tail -n 4 gpurun_out/${T}_pytest.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 profiles/sharded_time.py 1 2 4 > gpurun_out/${T}_sharded_time.txt 2> gpurun_out/${T}_sharded_time.err; echo "sharded_time rc=$?"
tail -n 3 gpurun_out/${T}_sharded_time.err
cat gpurun_out/${T}_sharded_time.txt
