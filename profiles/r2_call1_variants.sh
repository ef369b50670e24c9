# round 2, GPU call 1: timing of the k_rows variants compiled in round 1 (no parity run; parity is in the suite)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
for C in 0 1 2 4; do
  echo "== NVNL_ROWS_CONFIG=$C"
  NVNL_ROWS_CONFIG=$C timeout 300 python profiles/configs_api_time.py gpurun_out/r2c1_variants_cfg${C}.json 2>&1 | grep -E "cfg[2345]"
  NVNL_ROWS_CONFIG=$C timeout 100 python profiles/loop_cfg4.py 2>&1 | grep "flush=1 keep_out=0"
done
