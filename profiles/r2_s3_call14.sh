mkdir -p gpurun_out
T=r2s3c14
for v in nvalchemi-toolkit-ops_b200/csrc/libnvalchemi_nl_b200.so nvalchemi-toolkit-ops_b200/csrc/variants/lib_prev.so; do
  timeout 200 python profiles/variant_time.py $v 2>&1 | grep -E "ms" | tee -a gpurun_out/${T}_variants.txt
done
timeout 600 python -m pytest tests -m gpu -q --timeout 500 -p no:cacheprovider -x -k "single_cell or coo_paths or config3 or config5 or config4 or speculative or fused_sweep" > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?"
tail -n 3 gpurun_out/${T}_pytest.log
