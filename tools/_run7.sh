set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest7.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest7.log
timeout 600 python tools/e2e_profile.py > gpurun_out/r2_e2e_prof7.log 2>&1
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench7.log 2> gpurun_out/r2_bench7.err
grep -n "^FAILED\|passed\|failed\|^E  " gpurun_out/r2_pytest7.log | tail -20; cat gpurun_out/r2_bench7.log; head -30 gpurun_out/r2_e2e_prof7.log; tail -3 gpurun_out/r2_bench7.err
