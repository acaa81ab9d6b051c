timeout 600 python -m pytest tests/test_dense_gpu.py tests/test_runtime_gpu.py tests/test_rt_wide_deep_gpu.py -m gpu -q --timeout=600 -x -k "column_blocks or torch" > gpurun_out/r2ag_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2ag_pytest.log
tail -25 gpurun_out/r2ag_pytest.log
