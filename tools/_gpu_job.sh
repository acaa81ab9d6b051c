timeout 1500 python -m pytest tests -m gpu -q --timeout=900 > gpurun_out/r2ah_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2ah_pytest.log
tail -6 gpurun_out/r2ah_pytest.log
timeout 600 python bench.py --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/r2ah_bench1.json 2> gpurun_out/r2ah_bench1.err; echo "bench rc=$?"
python __graft_entry__.py --smoke 2>&1 | tail -1
