# NCCL tuning sweep of the chunked exchange at N GPUs (each setting = its own torchrun job; sharded_time.py prints ms per call)
N=${1:-2}
mkdir -p gpurun_out
T=r2s4nccl${N}
run() {
  name=$1; shift
  echo "== $name" >> gpurun_out/${T}.txt
  env "$@" timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 profiles/sharded_time.py 2 2>> gpurun_out/${T}.err | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l)['by_chunks']['2']
        print(d['ms_per_call_max_over_ranks'], d['phase_ms_rank0']['nccl_chunks_on_comm_stream'], d['agree'])
" >> gpurun_out/${T}.txt
}
run default X=1
run min_nchannels32 NCCL_MIN_NCHANNELS=32
run ctas64 NCCL_MIN_CTAS=64 NCCL_MAX_CTAS=64
run ctas32 NCCL_MIN_CTAS=32 NCCL_MAX_CTAS=32
run simple_ring NCCL_PROTO=Simple NCCL_ALGO=Ring
run nvls0 NCCL_NVLS_ENABLE=0
run buff8m NCCL_BUFFSIZE=8388608
run p2p_chunk NCCL_P2P_NET_CHUNKSIZE=1048576 NCCL_CHUNK_SIZE=1048576
NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,TUNING timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 profiles/sharded_time.py 2 2>&1 | grep -iE "channels|nvls|AllGather|Connected|comm 0x.*rank 0" | head -n 30 > gpurun_out/${T}_info.txt
cat gpurun_out/${T}.txt
