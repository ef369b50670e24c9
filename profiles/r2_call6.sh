mkdir -p gpurun_out
T=r2c6
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?"
tail -8 gpurun_out/${T}_pytest.log
for L in profiles/variants/lib_c5_b4_r44_o2048.so profiles/variants/lib_c5_b4_r44_o1024.so profiles/variants/lib_c5_b4_r44_o4096.so; do
  timeout 300 python profiles/variant_time.py $L 2>&1 | grep -E "parity|ms|Error|error" | tail -6
done
timeout 300 python profiles/variant_time.py nvalchemi-toolkit-ops_b200/csrc/libnvalchemi_nl_b200.so 2>&1 | grep -E "parity|ms|Error|error" | tail -6
