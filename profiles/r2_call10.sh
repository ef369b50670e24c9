mkdir -p gpurun_out
T=r2c10
timeout 300 python profiles/variant_time.py nvalchemi-toolkit-ops_b200/csrc/libnvalchemi_nl_b200.so 2>&1 | grep -E "parity|ms|Error|error" | tail -6
timeout 600 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider -x -k "coo_paths or config4 or prezero or speculative or known_answer or published or sharded or config5" > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/${T}_pytest.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-other-configs --no-sharded 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('ms_per_step', d['ms_per_step'], 'first', d['first_call_ms'], d['roofline']['stages_ms'], 'launches/step', d['gpu_launches_per_step'])"
