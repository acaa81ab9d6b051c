timeout 400 python -m pytest tests/test_unique_gpu.py tests/test_sparse_opt_gpu.py tests/test_peer_sharded_gpu.py tests/test_hash_gpu.py tests/test_gather_gpu.py tests/test_wide_deep_gpu.py tests/test_dynamic_embedding_gpu.py -m gpu -q -x --timeout=120 > gpurun_out/r2d_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2d_pytest.log
timeout 120 python tools/kbench.py unique > gpurun_out/r2d_kbench_onesweep.log 2>&1
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2d_unique_launches.csv python tools/kbench.py unique --eager > gpurun_out/r2d_ncu.log 2>&1
timeout 300 python bench.py --steps 20 --warmup 3 --no-extra-configs --no-cpu-baseline --layout interleaved > gpurun_out/r2d_bench_inter.json 2> gpurun_out/r2d_bench_inter.err
timeout 300 python bench.py --steps 20 --warmup 3 --no-extra-configs --no-cpu-baseline --layout split > gpurun_out/r2d_bench_split.json 2> gpurun_out/r2d_bench_split.err
tail -3 gpurun_out/r2d_pytest.log
