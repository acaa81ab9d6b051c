N=${N:-4}
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 100 --warmup 5 > gpurun_out/r2ae_bench$N.json 2> gpurun_out/r2ae_bench$N.err; echo "bench rc=$?"
