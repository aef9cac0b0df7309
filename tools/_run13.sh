set -x
timeout 600 python -m pytest tests/test_irc_gpu.py tests/test_full_size_gpu.py -m gpu -q -x 2>&1 | tail -3
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench13.log 2> gpurun_out/r2_bench13.err
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --paths-log2 21 > gpurun_out/r2_bench13_2p21.log 2>> gpurun_out/r2_bench13.err
python - <<'PY'
import json
for f in ("gpurun_out/r2_bench13.log","gpurun_out/r2_bench13_2p21.log"):
    d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, "%.4e"%d["value"], d["ms_per_step"], "e2e %.4e"%d["e2e"]["value"])
PY
