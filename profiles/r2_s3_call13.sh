mkdir -p gpurun_out
T=r2s3c13
for v in nvalchemi-toolkit-ops_b200/csrc/libnvalchemi_nl_b200.so nvalchemi-toolkit-ops_b200/csrc/variants/lib_*.so; do
  timeout 200 python profiles/variant_time.py $v 2>&1 | grep -E "n=1000000" | tee -a gpurun_out/${T}_variants.txt
done
