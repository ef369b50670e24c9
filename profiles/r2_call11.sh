mkdir -p gpurun_out
T=r2c11
timeout 300 python profiles/variant_time.py nvalchemi-toolkit-ops_b200/csrc/libnvalchemi_nl_b200.so 2>&1 | grep -E "parity|ms|Error|error" | tail -6
timeout 600 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider -x -k "coo_paths or config4 or prezero or speculative or known_answer or published or sharded" > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/${T}_pytest.log
for spec in 1 0; do
NVNL_SPEC=$spec timeout 300 python - <<'PY'
import os, sys, json, subprocess
sys.path.insert(0, 'nvalchemi-toolkit-ops_b200')
from nvalchemiops_b200 import config
config.speculative_fill = bool(int(os.environ['NVNL_SPEC']))
sys.argv = ['bench.py', '--steps', '12', '--warmup', '3', '--no-cpu-baseline', '--no-other-configs', '--no-sharded']
import io, contextlib, runpy
buf = io.StringIO()
with contextlib.redirect_stdout(buf):
    try:
        runpy.run_path('bench.py', run_name='__main__')
    except SystemExit:
        pass
d = json.loads(buf.getvalue().strip().splitlines()[-1])
print('speculative', os.environ['NVNL_SPEC'], 'ms_per_step', round(d['ms_per_step'], 4), 'first', round(d['first_call_ms'], 3), {k: round(v, 4) for k, v in d['roofline']['stages_ms'].items()}, 'launches/step', d['gpu_launches_per_step'])
PY
done
for C in 3 5; do
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/${T}_launches_cfg${C}.csv python profiles/cfg_calls.py $C 3 > gpurun_out/${T}_cfg${C}.log 2>&1
python profiles/launch_list.py gpurun_out/${T}_launches_cfg${C}.csv
timeout 100 python profiles/cfg_calls.py $C 4 | tail -2
done
