timeout 900 python -m pytest tests -m gpu -q --timeout=600 > gpurun_out/r2e_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2e_pytest.log
timeout 120 python tools/kbench.py unique > gpurun_out/r2e_kbench_onesweep.log 2>&1
MREC_UNIQUE_LSD=1 timeout 120 python tools/kbench.py unique > gpurun_out/r2e_kbench_lsd.log 2>&1
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2e_unique_launches.csv python tools/kbench.py unique --eager > gpurun_out/r2e_ncu.log 2>&1
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err; echo "bench rc=$?" >> gpurun_out/r2e_bench.err
tail -5 gpurun_out/r2e_pytest.log
