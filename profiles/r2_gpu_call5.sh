set -x
mkdir -p gpurun_out
T=${TAG:-c8}
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider -x > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/${T}_pytest.log
timeout 400 python bench.py --steps 30 --warmup 5 > gpurun_out/${T}_bench_rows.json 2> gpurun_out/${T}_bench_rows.err; echo "bench rc=$?"; cat gpurun_out/${T}_bench_rows.json
timeout 300 python profiles/configs_api_time.py gpurun_out/${T}_configs.json > gpurun_out/${T}_configs.log 2>&1; tail -14 gpurun_out/${T}_configs.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 90 --csv --log-file gpurun_out/${T}_launches.csv python profiles/cfg4_calls.py 3 rows > gpurun_out/${T}_ncu_launch.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'k_rows|k_zero' -s 3 -c 3 -o gpurun_out/${T}_rows_full -f python profiles/cfg4_calls.py 2 rows > gpurun_out/${T}_ncu_full.log 2>&1; echo "ncu rc=$?"
