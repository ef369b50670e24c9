mkdir -p gpurun_out
T=r2s3g
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider > gpurun_out/${T}_pytest.log 2>&1; echo "pytest full rc=$?"
tail -n 4 gpurun_out/${T}_pytest.log
timeout 900 python bench.py --steps 30 --warmup 5 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"; tail -n 2 gpurun_out/${T}_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2s3g_bench.json'))
print('ms_per_step', d['ms_per_step'], 'first', d['first_call_ms'], 'frac', d['roofline']['frac'], d['roofline']['stages_ms'])
print({k:v['ms_per_call'] for k,v in d['other_configs'].items()}, d['other_configs']['config2']['max_neighbors_160']['ms_per_call'])
print({k:(v['fused_ms'], v['list_pipeline_ms']) for k,v in d['fused_consumer'].items() if k.startswith('alpha')})
print('e2e', d['e2e']['ms_per_step'], 'cpu', d['cpu_baseline']['value'], 'launches/step', d['gpu_launches_per_step'], d['clocks'])
PY
for C in 3; do
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/${T}_launches_cfg${C}.csv python profiles/cfg_calls.py $C 3 > gpurun_out/${T}_cfg${C}.log 2>&1
echo "== config $C"; python profiles/launch_list.py gpurun_out/${T}_launches_cfg${C}.csv
done
timeout 400 compute-sanitizer --tool racecheck --print-limit 5 python -m pytest tests/test_gpu_parity.py -q --timeout 380 -p no:cacheprovider -k "coo_paths_medium_box" > gpurun_out/${T}_racecheck.log 2>&1; echo "racecheck rc=$?"
grep -n "RACECHECK SUMMARY\|ERROR SUMMARY\|passed\|failed\|hazard" gpurun_out/${T}_racecheck.log | tail -n 5
timeout 400 compute-sanitizer --tool initcheck --print-limit 5 python -m pytest tests/test_gpu_parity.py -q --timeout 380 -p no:cacheprovider -k "coo_paths_medium_box or single_cell" > gpurun_out/${T}_initcheck.log 2>&1; echo "initcheck rc=$?"
grep -n "ERROR SUMMARY\|passed\|failed\|Uninitialized" gpurun_out/${T}_initcheck.log | tail -n 5
