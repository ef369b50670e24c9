mkdir -p gpurun_out
T=r2s4c4
timeout 120 python -m pytest tests -m gpu -q --timeout 100 -p no:cacheprovider -x -k "not config4 and not config5 and not config3_full and not real_warp and not nccl and not published_fcc and not config2" > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?"
tail -n 6 gpurun_out/${T}_pytest.log
