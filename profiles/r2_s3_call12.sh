mkdir -p gpurun_out
T=r2s3c12
for v in nvalchemi-toolkit-ops_b200/csrc/libnvalchemi_nl_b200.so nvalchemi-toolkit-ops_b200/csrc/variants/lib_*.so; do
  timeout 200 python profiles/variant_time.py $v 2>&1 | grep -E "parity|ms|Error|error" | tee -a gpurun_out/${T}_variants.txt
done
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider -x -k "coo_paths or config4 or config5 or config3 or prezero or speculative or known_answer or sharded or single_cell or overflow" > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?"
tail -n 4 gpurun_out/${T}_pytest.log
