set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_irc_gpu.py tests/test_full_size_gpu.py -m gpu -q -x 2>&1 | tail -5 > gpurun_out/r2_pytest4.log
timeout 900 python tools/tune_irc.py run 4x3 4x3,MCRE_CVA_PF=0 8x2,MCRE_CVA_PF=0 6x2 5x3 > gpurun_out/r2_tune4.log 2>&1
MCRE_LIB_PATH=montecarlo-risk-engine_b200/variants/libmcre_b200_4x3.so timeout 600 ncu --set full --clock-control none --import-source on -k regex:irc_cva_kernel -s 1 -c 1 -f -o gpurun_out/r2_cva_v3_4x3 python bench.py --steps 2 --warmup 1 --paths-log2 22 --presim-log2 18 --no-cpu-baseline --no-e2e > gpurun_out/r2_ncu_v3.log 2>&1
cat gpurun_out/r2_pytest4.log gpurun_out/r2_tune4.log
