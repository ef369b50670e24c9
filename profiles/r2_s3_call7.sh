mkdir -p gpurun_out
T=r2s3c7
timeout 600 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider -x -k "single_cell or coo_paths or config3 or known_answer" > gpurun_out/${T}_pytest_a.log 2>&1; echo "pytest subset rc=$?"
tail -n 15 gpurun_out/${T}_pytest_a.log
timeout 200 python profiles/variant_time.py nvalchemi-toolkit-ops_b200/csrc/libnvalchemi_nl_b200.so 2>&1 | grep -E "parity|ms|Error|error" | tee -a gpurun_out/${T}_variants.txt
for C in 3 4 5; do
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/${T}_launches_cfg${C}.csv python profiles/cfg_calls.py $C 3 > gpurun_out/${T}_cfg${C}.log 2>&1
python profiles/launch_list.py gpurun_out/${T}_launches_cfg${C}.csv
done
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider > gpurun_out/${T}_pytest.log 2>&1; echo "pytest full rc=$?"
tail -n 6 gpurun_out/${T}_pytest.log
