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
echo "== experimental: output kernel before the size sync (config.speculative_fill)"
NVNL_EXPERIMENTAL=1 timeout 300 python -m pytest tests -m gpu -q -x -p no:cacheprovider -k experimental 2>&1 | tail -1
timeout 100 python - <<'PY'
import sys, torch
sys.path[:0] = ['.', 'nvalchemi-toolkit-ops_b200', 'tests', 'oracle']
from systems import bench_box
from nvalchemiops_b200 import config
from nvalchemiops_b200.neighborlist import neighbor_list
pos, cell, pbc = [t.to('cuda:0') for t in bench_box(1_000_000, seed=4)]
for spec in (False, True):
    config.speculative_fill = spec
    for _ in range(4): out = neighbor_list(pos, 6.0, cell=cell, pbc=pbc, return_neighbor_list=True)
    del out; torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(20)]
    for a, b in ev:
        a.record(); out = neighbor_list(pos, 6.0, cell=cell, pbc=pbc, return_neighbor_list=True); b.record(); del out
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in ev)
    print('speculative_fill=%d median %.3f ms' % (spec, ts[10]))
PY
