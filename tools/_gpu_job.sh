timeout 600 python -m pytest tests/test_peer_step_gpu.py tests/test_peer_sharded_gpu.py -q -x --timeout=600 > gpurun_out/r2ad_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2ad_pytest.log
tail -3 gpurun_out/r2ad_pytest.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 100 --warmup 5 > gpurun_out/r2ad_bench2.json 2> gpurun_out/r2ad_bench2.err; echo "bench rc=$?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 tools/timeline_sharded.py > gpurun_out/r2ad_timeline_wd_n2.txt 2> gpurun_out/r2ad_timeline_wd_n2.err; echo "rc=$?"
rm -f gpurun_out/timeline_*.json
