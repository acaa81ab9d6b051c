timeout 900 python -m pytest tests/test_peer_step_gpu.py tests/test_peer_sharded_gpu.py tests/test_sharded_gpu.py tests/test_multitable_sharded_gpu.py tests/test_dense_gpu.py tests/test_wide_deep_gpu.py -q -x --timeout=600 > gpurun_out/r2r_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2r_pytest.log
tail -5 gpurun_out/r2r_pytest.log
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 100 --warmup 5 > gpurun_out/r2r_bench2.json 2> gpurun_out/r2r_bench2.err; echo "bench rc=$?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 tools/timeline_sharded.py --c5 > gpurun_out/r2r_timeline_c5_n2.txt 2> gpurun_out/r2r_timeline_c5_n2.err; echo "rc=$?"
rm -f gpurun_out/timeline_*.json
