set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest6.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest6.log
timeout 600 python tools/e2e_profile.py > gpurun_out/r2_e2e_prof6.log 2>&1
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench6.log 2> gpurun_out/r2_bench6.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2_launches6.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r2_ncu_launch6.log 2>&1
grep -n "^FAILED\|passed\|failed" gpurun_out/r2_pytest6.log | tail; cat gpurun_out/r2_bench6.log; head -45 gpurun_out/r2_e2e_prof6.log
