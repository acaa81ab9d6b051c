timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/diag_parity.py > gpurun_out/r2i_diag.log 2>&1; echo "rc=$?" >> gpurun_out/r2i_diag.log
grep -v "^W\|^\[rank" gpurun_out/r2i_diag.log | tail -20
