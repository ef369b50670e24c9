mkdir -p gpurun_out
T=r2s3c2
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_rows --launch-skip 4 --launch-count 2 -o gpurun_out/${T}_cfg4 -f python profiles/cfg_calls.py 4 3 > gpurun_out/${T}_cfg4.log 2>&1; echo "cfg4 rc=$?"
timeout 600 ncu --set full --import-source on --clock-control none -k "regex:k_rows<" --launch-skip 2 --launch-count 1 -o gpurun_out/${T}_cfg5 -f python profiles/cfg_calls.py 5 3 > gpurun_out/${T}_cfg5.log 2>&1; echo "cfg5 rc=$?"
timeout 600 ncu --set full --import-source on --clock-control none -k "regex:k_rows<|k_sweep" --launch-skip 6 --launch-count 3 -o gpurun_out/${T}_cfg3 -f python profiles/cfg_calls.py 3 3 > gpurun_out/${T}_cfg3.log 2>&1; echo "cfg3 rc=$?"
ls -la gpurun_out/*.ncu-rep
