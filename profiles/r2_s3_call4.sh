mkdir -p gpurun_out
T=r2s3c4
timeout 200 python profiles/variant_time.py nvalchemi-toolkit-ops_b200/csrc/libnvalchemi_nl_b200.so 2>&1 | grep -E "parity|ms|Error|error" | tee -a gpurun_out/${T}_variants.txt
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider -x -k "coo_paths or config5 or config3 or prezero or speculative or known_answer or sharded or batch" > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?"
tail -n 4 gpurun_out/${T}_pytest.log
for C in all big small3 small2 small1; do
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/${T}_cfg3_${C}.csv python profiles/cfg3_split.py $C 3 > gpurun_out/${T}_cfg3_${C}.log 2>&1
echo "== $C"; tail -n 1 gpurun_out/${T}_cfg3_${C}.log
python profiles/launch_list.py gpurun_out/${T}_cfg3_${C}.csv | grep -E "k_rows|k_sweep|sum"
done
