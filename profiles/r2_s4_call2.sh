mkdir -p gpurun_out
T=r2s4c2
timeout 400 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider -k "chunked_exchange or sharded or known_answer_counts or coo_paths_medium" > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?"
tail -n 15 gpurun_out/${T}_pytest.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 2
