set -x
mkdir -p gpurun_out
T=${TAG:-fin}
timeout 300 python bench.py --steps 30 --warmup 5 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"
timeout 120 python profiles/configs_api_time.py gpurun_out/${T}_configs.json > gpurun_out/${T}_configs.log 2>&1
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 90 --csv --log-file gpurun_out/${T}_launches.csv python profiles/cfg4_calls.py 3 rows > gpurun_out/${T}_ncu_launch.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'k_rows' -s 2 -c 2 -o gpurun_out/${T}_rows_full -f python profiles/cfg4_calls.py 2 rows > gpurun_out/${T}_ncu_full.log 2>&1; echo "ncu rc=$?"
