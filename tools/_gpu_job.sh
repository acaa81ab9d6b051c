timeout 300 python -m pytest tests/test_peer_sharded_gpu.py -m gpu -q --timeout=300 -x -k "one_step_ahead" > gpurun_out/r2aj_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2aj_pytest.log
tail -15 gpurun_out/r2aj_pytest.log
