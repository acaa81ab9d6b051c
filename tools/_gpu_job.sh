timeout 900 python bench.py > gpurun_out/r2z_bench1.json 2> gpurun_out/r2z_bench1.err; echo "bench rc=$?"
tail -3 gpurun_out/r2z_bench1.err
