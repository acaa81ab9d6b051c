N=${N:-2}
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 100 --warmup 5 > gpurun_out/r2l_bench$N.json 2> gpurun_out/r2l_bench$N.err; echo "bench rc=$?" >> gpurun_out/r2l_bench$N.err
grep -v "^W\|^\[rank" gpurun_out/r2l_bench$N.err | tail -5
