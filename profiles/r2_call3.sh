set -x
mkdir -p gpurun_out
T=${TAG:-r2c3}
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/${T}_launches.csv python profiles/cfg4_calls.py 3 rows > gpurun_out/${T}_ncu_launch.log 2>&1
grep -E "k_rows|k_scan|k_hash|k_scatter" gpurun_out/${T}_launches.csv | awk -F'","' '{print $5, $NF}' | cut -c1-60,200- | tail -12
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_rows' -s 2 -c 2 -o gpurun_out/${T}_rows_full -f python profiles/cfg4_calls.py 2 rows > gpurun_out/${T}_ncu_full.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out/${T}_rows_full.ncu-rep
