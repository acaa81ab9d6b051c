timeout 1500 python -m pytest tests -m gpu -q --timeout=900 -x > gpurun_out/r2v_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2v_pytest.log
tail -5 gpurun_out/r2v_pytest.log
