timeout 600 python -m pytest tests/test_rt_wide_deep_gpu.py -m gpu -q --timeout=600 -x > gpurun_out/r2ai_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2ai_pytest.log
tail -15 gpurun_out/r2ai_pytest.log
timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/r2ai_bench1.json 2> gpurun_out/r2ai_bench1.err; echo "bench rc=$?"
