set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest11.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest11.log
grep -n "^FAILED\|passed\|failed\|^E  " gpurun_out/r2_pytest11.log | tail -20
timeout 900 python tools/run_configs.py 1 2 2o 4 5 5g > gpurun_out/r2_configs11.jsonl 2> gpurun_out/r2_configs11.err
python - <<'PY'
import json
for l in open("gpurun_out/r2_configs11.jsonl"):
    d=json.loads(l); print(d["config"], "%.1f ms"%(d["seconds"]*1e3), "%.3e"%d["path_steps_per_s"], d["launches"], d["timings"])
PY
tail -n 5 gpurun_out/r2_configs11.err
