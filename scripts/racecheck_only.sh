SELR='window_fp64_vs_clean_oracle and not 300 or persistent_batch'
timeout 1500 compute-sanitizer --tool racecheck --racecheck-report all --error-exitcode 9 --log-file gpurun_out/r02_racecheck.log python -m pytest tests/test_engine_gpu.py -m gpu -q -x -k "$SELR" > gpurun_out/r02_racecheck_pytest.log 2>&1
echo "racecheck exit $?" >> gpurun_out/r02_racecheck_pytest.log
tail -n 3 gpurun_out/r02_racecheck_pytest.log; tail -n 3 gpurun_out/r02_racecheck.log
