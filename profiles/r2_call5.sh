mkdir -p gpurun_out
T=r2c5
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?"
tail -8 gpurun_out/${T}_pytest.log
timeout 300 python profiles/variant_time.py nvalchemi-toolkit-ops_b200/csrc/libnvalchemi_nl_b200.so 2>&1 | grep -E "parity|ms|Error|error" | tail -8
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_rows' -s 2 -c 2 -o gpurun_out/${T}_rows_full -f python profiles/cfg4_calls.py 2 rows > gpurun_out/${T}_ncu_full.log 2>&1; echo "ncu rc=$?"
