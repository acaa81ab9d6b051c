timeout 1500 python -m pytest tests -m gpu -q --timeout=900 -x > gpurun_out/r2aa_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2aa_pytest.log
tail -4 gpurun_out/r2aa_pytest.log
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extra-configs > gpurun_out/r2aa_plain.json 2> gpurun_out/r2aa_plain.err; echo "plain rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r2_launches_bench.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extra-configs > gpurun_out/r2aa_ncu.log 2>&1; echo "ncu rc=$?"
python __graft_entry__.py --smoke 2>&1 | tail -2
