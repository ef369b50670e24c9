mkdir -p gpurun_out
T=r2s4c1
timeout 600 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider -k "public_naive or dispatcher or dual_cutoff or rebuild_detection or empty_and_zero" > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?"
tail -n 30 gpurun_out/${T}_pytest.log
