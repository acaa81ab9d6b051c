timeout 1500 python -m pytest tests -m gpu -q --timeout=900 -x > gpurun_out/r2x_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2x_pytest.log
tail -5 gpurun_out/r2x_pytest.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 100 --warmup 5 > gpurun_out/r2x_bench2.json 2> gpurun_out/r2x_bench2.err; echo "bench rc=$?"
