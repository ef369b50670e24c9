mkdir -p gpurun_out
T=r2s4c3
timeout 400 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider -k "chunked_exchange or sharded" > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?"
tail -n 15 gpurun_out/${T}_pytest.log
