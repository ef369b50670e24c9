set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
timeout 300 python __graft_entry__.py smoke > gpurun_out/c1_smoke.log 2>&1; echo "smoke rc=$?"
tail -3 gpurun_out/c1_smoke.log
timeout 300 python profiles/cfg4_calls.py 4 rows > gpurun_out/c1_rows.log 2>&1; tail -4 gpurun_out/c1_rows.log
timeout 300 python profiles/cfg4_calls.py 4 masks > gpurun_out/c1_masks.log 2>&1; tail -4 gpurun_out/c1_masks.log
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/c1_pytest.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/c1_pytest.log
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/c1_bench_rows.json 2> gpurun_out/c1_bench_rows.err; echo "bench rc=$?"; cat gpurun_out/c1_bench_rows.json
timeout 300 python bench.py --steps 20 --warmup 5 --coo-path masks --no-cpu-baseline > gpurun_out/c1_bench_masks.json 2> gpurun_out/c1_bench_masks.err; cat gpurun_out/c1_bench_masks.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/c1_launches.csv python profiles/cfg4_calls.py 3 rows > gpurun_out/c1_ncu_launch.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'k_rows' -s 2 -c 2 -o gpurun_out/c1_rows_full -f python profiles/cfg4_calls.py 2 rows > gpurun_out/c1_ncu_full.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out
