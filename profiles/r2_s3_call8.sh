mkdir -p gpurun_out
T=r2s3c8
timeout 900 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_parity.py -q --timeout 800 -p no:cacheprovider -x -k "known_answer_counts or speculative_fill" > gpurun_out/${T}_memcheck.log 2>&1; echo "memcheck rc=$?"
grep -n "Invalid\|at nvnl\|at .*cuh\|ERROR SUMMARY\|=========     at" gpurun_out/${T}_memcheck.log | head -30
tail -n 5 gpurun_out/${T}_memcheck.log
