mkdir -p gpurun_out
T=r2s3h
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider > gpurun_out/${T}_pytest.log 2>&1; echo "pytest full rc=$?"
tail -n 3 gpurun_out/${T}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 1
timeout 900 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"; tail -n 2 gpurun_out/${T}_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2s3h_bench.json'))
print('ms_per_step', d['ms_per_step'], 'first', d['first_call_ms'], 'frac', d['roofline']['frac'], d['roofline']['stages_ms'])
print({k:v['ms_per_call'] for k,v in d['other_configs'].items()}, d['other_configs']['config2']['max_neighbors_160']['ms_per_call'])
print({k:(v['fused_ms'], v['list_pipeline_ms']) for k,v in d['fused_consumer'].items() if k.startswith('alpha')})
print('e2e', d['e2e']['ms_per_step'], 'cpu', d['cpu_baseline']['value'], 'launches/step', d['gpu_launches_per_step'], d['clocks'])
PY
