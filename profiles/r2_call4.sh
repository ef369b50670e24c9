mkdir -p gpurun_out
for L in profiles/variants/lib_c5_b4_r44.so profiles/variants/lib_c8_b3_r60.so profiles/variants/lib_c7_b3_r60.so profiles/variants/lib_c4_b5_r36.so; do
  timeout 300 python profiles/variant_time.py $L 2>&1 | grep -E "parity|ms|Error|error" | tail -8
done
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider -x > gpurun_out/r2c4_pytest.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r2c4_pytest.log
