set -x
mkdir -p gpurun_out
for n in 8 4 2; do
for sc in weak strong; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 10 --warmup 3 --scaling $sc > gpurun_out/r2f_bench_${n}gpu_$sc.log 2> gpurun_out/r2f_bench_${n}gpu_$sc.err
done
done
python tools/multi_gpu_check.py --out gpurun_out/r2f_mg1.json > gpurun_out/r2f_mg1.log 2>&1
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29522 tools/multi_gpu_check.py --out gpurun_out/r2f_mg2.json > gpurun_out/r2f_mg2.log 2>&1
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 tools/multi_gpu_check.py --out gpurun_out/r2f_mg8.json > gpurun_out/r2f_mg8.log 2>&1
python tools/multi_gpu_check.py --compare gpurun_out/r2f_mg1.json gpurun_out/r2f_mg2.json > gpurun_out/r2f_mg_compare_2gpu.log 2>&1
python tools/multi_gpu_check.py --compare gpurun_out/r2f_mg1.json gpurun_out/r2f_mg8.json > gpurun_out/r2f_mg_compare_8gpu.log 2>&1
cat gpurun_out/r2f_mg_compare_2gpu.log gpurun_out/r2f_mg_compare_8gpu.log
python - <<'PY'
import json
for n in (2,4,8):
  for f in ("weak","strong"):
    try:
        d=json.loads(open(f"gpurun_out/r2f_bench_{n}gpu_{f}.log").read().strip().splitlines()[-1])
        print(n, f, "%.4e"%d["value"], d["ms_per_step"], "e2e %.4e"%d["e2e"]["value"])
    except Exception as e: print(n,f,"failed",e)
PY
