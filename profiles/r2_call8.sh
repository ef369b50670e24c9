mkdir -p gpurun_out
T=r2c8
for L in nvalchemi-toolkit-ops_b200/csrc/libnvalchemi_nl_b200.so profiles/variants/lib_out3cta.so; do
  timeout 300 python profiles/variant_time.py $L 2>&1 | grep -E "parity|ms|Error|error" | tail -6
done
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?"
tail -6 gpurun_out/${T}_pytest.log
