mkdir -p gpurun_out
T=r2s3c10
timeout 900 python -m pytest tests/test_gpu_pair_consumer.py -q --timeout 600 -p no:cacheprovider > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?"
tail -n 25 gpurun_out/${T}_pytest.log | cut -c1-200
timeout 300 python profiles/pair_time.py 1000000 0.3 2>&1 | tail -n 3 | tee gpurun_out/${T}_pair_time.txt
timeout 300 python profiles/pair_time.py 1000000 0.0 2>&1 | tail -n 1 | tee -a gpurun_out/${T}_pair_time.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 2
