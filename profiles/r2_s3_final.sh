mkdir -p gpurun_out
T=r2s3f
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider > gpurun_out/${T}_pytest.log 2>&1; echo "pytest full rc=$?"
tail -n 4 gpurun_out/${T}_pytest.log
timeout 900 python bench.py --steps 30 --warmup 5 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"; tail -n 2 gpurun_out/${T}_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2s3f_bench.json'))
print('ms_per_step', d['ms_per_step'], 'first', d['first_call_ms'], 'frac', d['roofline']['frac'], d['roofline']['stages_ms'])
print({k:v['ms_per_call'] for k,v in d['other_configs'].items()}, d['other_configs']['config2']['max_neighbors_160']['ms_per_call'])
print(d.get('fused_consumer'))
print('e2e', d['e2e']['ms_per_step'], 'cpu', d['cpu_baseline']['value'], 'launches/step', d['gpu_launches_per_step'], d['clocks'])
PY
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${T}_bench_ref.json 2>/dev/null; echo "ref rc=$?"; head -c 600 gpurun_out/${T}_bench_ref.json; echo
timeout 600 ncu --set full --import-source on --clock-control none -k "regex:^k_rows" --launch-skip 4 --launch-count 2 -o gpurun_out/${T}_cfg4 -f python profiles/cfg_calls.py 4 3 > gpurun_out/${T}_cfg4.log 2>&1; echo "ncu rc=$?"
for C in 3 4 5 2; do
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/${T}_launches_cfg${C}.csv python profiles/cfg_calls.py $C 3 > gpurun_out/${T}_cfg${C}.log 2>&1
echo "== config $C"; python profiles/launch_list.py gpurun_out/${T}_launches_cfg${C}.csv
done
timeout 900 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_parity.py tests/test_gpu_pair_consumer.py -q --timeout 800 -p no:cacheprovider -k "coo_paths or single_cell or known_answer or fused_sweep_matches or list_consumer" > gpurun_out/${T}_memcheck.log 2>&1; echo "memcheck rc=$?"
grep -n "ERROR SUMMARY\|passed\|failed" gpurun_out/${T}_memcheck.log | tail -n 4
