mkdir -p gpurun_out
T=r2s3c1
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?"
tail -6 gpurun_out/${T}_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"
cat gpurun_out/${T}_bench.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/${T}_launches_cfg4.csv python profiles/cfg_calls.py 4 3 > gpurun_out/${T}_cfg4.log 2>&1
python profiles/launch_list.py gpurun_out/${T}_launches_cfg4.csv
for C in 3 5 2; do
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/${T}_launches_cfg${C}.csv python profiles/cfg_calls.py $C 3 > gpurun_out/${T}_cfg${C}.log 2>&1
python profiles/launch_list.py gpurun_out/${T}_launches_cfg${C}.csv
done
