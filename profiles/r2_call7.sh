mkdir -p gpurun_out
T=r2c7
timeout 300 python profiles/variant_time.py nvalchemi-toolkit-ops_b200/csrc/libnvalchemi_nl_b200.so 2>&1 | grep -E "parity|ms|Error|error" | tail -6
timeout 600 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider -x -k "coo_paths or config4 or prezero or estimate or known_answer or published" > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/${T}_pytest.log
cat > /tmp/noprezero.py <<'PY'
import sys, runpy
sys.argv = ['cfg4_calls.py', '3', 'rows']
sys.path.insert(0, 'nvalchemi-toolkit-ops_b200')
from nvalchemiops_b200 import config
config.prezero_shifts = False
runpy.run_path('profiles/cfg4_calls.py', run_name='__main__')
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_rows' -s 2 -c 2 -o gpurun_out/${T}_rows_noprezero -f python /tmp/noprezero.py > gpurun_out/${T}_ncu_a.log 2>&1; echo "ncu rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_rows' -s 2 -c 2 -o gpurun_out/${T}_rows_prezero -f python profiles/cfg4_calls.py 3 rows > gpurun_out/${T}_ncu_b.log 2>&1; echo "ncu rc=$?"
timeout 300 python bench.py --steps 10 --warmup 3 --cpu-budget-s 4 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/${T}_bench.err; cat gpurun_out/${T}_bench.json | head -c 3000
