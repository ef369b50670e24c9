# First GPU call of the next round: parity + timing of the compiled-but-unmeasured k_rows variants.
#   NVNL_ROWS_CONFIG bit 0 = 2-stage ring, 4 CTAs/SM (36 instead of 27 warps/SM, 56 registers)
#                    bit 1 = 128-byte aligned, padded temporary rows (k_rows_out reads whole lines)
#                    bit 2 = chunks of boundary cells inside one image segment swept in groups of four (config 5, small boxes)
# usage: gpurun --timeout 1500 -- 'bash profiles/next_round_variants.sh'
mkdir -p gpurun_out
for C in 0 1 2 4 3 5 6 7; do
  echo "== NVNL_ROWS_CONFIG=$C"
  NVNL_ROWS_CONFIG=$C timeout 400 python -m pytest tests -m gpu -q -x -p no:cacheprovider \
      -k "coo_paths or prezero or overflow or batch or sharded or known_answer or small_systems or random_geometry or half_fill" 2>&1 | tail -1
  NVNL_ROWS_CONFIG=$C timeout 200 python profiles/configs_api_time.py gpurun_out/variants_cfg${C}.json 2>&1 | grep -E "cfg[2345].*rows"
  NVNL_ROWS_CONFIG=$C timeout 100 python profiles/loop_cfg4.py 2>&1 | grep "prezero=1 flush=1 keep_out=0"
done
