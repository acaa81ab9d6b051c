timeout 170 python -m pytest tests -m gpu -q --timeout=160 -x > gpurun_out/r2ak_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2ak_pytest.log
tail -4 gpurun_out/r2ak_pytest.log
