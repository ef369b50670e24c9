mkdir -p gpurun_out
T=r2s3c11
for v in nvalchemi-toolkit-ops_b200/csrc/libnvalchemi_nl_b200.so nvalchemi-toolkit-ops_b200/csrc/variants/lib_*.so; do
  timeout 200 python profiles/variant_time.py $v 2>&1 | grep -E "parity|ms|Error|error" | tee -a gpurun_out/${T}_variants.txt
done
timeout 600 ncu --set full --import-source on --clock-control none -k "regex:^k_rows" --launch-skip 6 --launch-count 3 -o gpurun_out/${T}_cfg3 -f python profiles/cfg_calls.py 3 3 > gpurun_out/${T}_cfg3.log 2>&1; echo "cfg3 rc=$?"
