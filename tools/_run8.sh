set -x
mkdir -p gpurun_out
python tools/multi_gpu_check.py --out gpurun_out/r2_mg1.json > gpurun_out/r2_mg1.log 2>&1
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/multi_gpu_check.py --out gpurun_out/r2_mg2.json > gpurun_out/r2_mg2.log 2>&1
python tools/multi_gpu_check.py --compare gpurun_out/r2_mg1.json gpurun_out/r2_mg2.json > gpurun_out/r2_mg_compare_2gpu.log 2>&1
cat gpurun_out/r2_mg_compare_2gpu.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2_bench_2gpu_weak.log 2> gpurun_out/r2_bench_2gpu_weak.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 5 --warmup 3 --scaling strong > gpurun_out/r2_bench_2gpu_strong.log 2> gpurun_out/r2_bench_2gpu_strong.err
python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_1gpu_b.log 2>/dev/null
cat gpurun_out/r2_bench_2gpu_weak.log gpurun_out/r2_bench_2gpu_strong.log gpurun_out/r2_bench_1gpu_b.log | cut -c1-900; tail -3 gpurun_out/r2_mg2.log gpurun_out/r2_bench_2gpu_weak.err
