mkdir -p gpurun_out
T=r2c9
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_rows_out' -s 1 -c 1 -o gpurun_out/${T}_out_ldg -f python profiles/cfg4_calls.py 3 rows > gpurun_out/${T}_ncu.log 2>&1; echo "ncu rc=$?"
