mkdir -p gpurun_out
T=r2s3c5
timeout 200 python profiles/variant_time.py nvalchemi-toolkit-ops_b200/csrc/libnvalchemi_nl_b200.so 2>&1 | grep -E "parity|ms|Error|error" | tee -a gpurun_out/${T}_variants.txt
NVNL_NO_PDL=1 timeout 200 python profiles/variant_time.py nvalchemi-toolkit-ops_b200/csrc/libnvalchemi_nl_b200.so 2>&1 | grep -E "parity|ms|Error|error" | sed 's/^/NO_PDL /' | tee -a gpurun_out/${T}_variants.txt
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider -x -k "coo_paths or config5 or config3 or prezero or speculative or known_answer or sharded or batch or matrix or split" > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?"
tail -n 4 gpurun_out/${T}_pytest.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-sharded > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2s3c5_bench.json'))
print('ms_per_step', d['ms_per_step'], 'first', d['first_call_ms'], d['roofline']['stages_ms'], {k:v['ms_per_call'] for k,v in d['other_configs'].items()})
PY
for C in 4 5; do
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/${T}_launches_cfg${C}.csv python profiles/cfg_calls.py $C 3 > gpurun_out/${T}_cfg${C}.log 2>&1
python profiles/launch_list.py gpurun_out/${T}_launches_cfg${C}.csv
done
