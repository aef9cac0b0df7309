set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest5.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest5.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench5.log 2> gpurun_out/r2_bench5.err
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_benchref5.log 2> gpurun_out/r2_benchref5.err
nproc; tail -5 gpurun_out/r2_pytest5.log; cat gpurun_out/r2_bench5.log gpurun_out/r2_benchref5.log; tail -5 gpurun_out/r2_bench5.err gpurun_out/r2_benchref5.err
