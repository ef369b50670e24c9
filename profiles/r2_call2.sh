set -x
mkdir -p gpurun_out
T=${TAG:-r2c2}
timeout 600 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider -x -k "coo_paths or known_answer" > gpurun_out/${T}_pytest_a.log 2>&1; echo "pytest-a rc=$?"
tail -15 gpurun_out/${T}_pytest_a.log
timeout 200 python profiles/stage_time.py 1000000 2>&1 | tail -5
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?"
tail -30 gpurun_out/${T}_pytest.log
