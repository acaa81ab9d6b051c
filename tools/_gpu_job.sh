# the round's standard validation job:  gpurun --timeout 1800 -- 'bash tools/_gpu_job.sh'
timeout 1500 python -m pytest tests -m gpu -q --timeout=900 -x > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 100 --warmup 5 > gpurun_out/bench1.json 2> gpurun_out/bench1.err; echo "bench rc=$?"
python __graft_entry__.py --smoke 2>&1 | tail -1
