set -x
export MCRE_BENCH_DEBUG=1
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r2_bench_2gpu_weak.log 2> gpurun_out/r2_bench_2gpu_weak.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 3 --warmup 3 --scaling strong > gpurun_out/r2_bench_2gpu_strong.log 2> gpurun_out/r2_bench_2gpu_strong.err
grep "rank" gpurun_out/r2_bench_2gpu_weak.err gpurun_out/r2_bench_2gpu_strong.err
python - <<'PY'
import json
for f in ("weak","strong"):
    d=json.loads(open(f"gpurun_out/r2_bench_2gpu_{f}.log").read().strip().splitlines()[-1])
    print(f, d["value"], d["e2e"]["value"])
PY
