timeout 900 python -m pytest tests/test_irc_gpu.py tests/test_fuzz_rates_gpu.py tests/test_full_size_gpu.py -m gpu -q 2>&1 | tail -3
timeout 900 python tools/run_configs.py 2 2o > gpurun_out/r2_configs20.jsonl 2> gpurun_out/r2_configs20.err
python - <<'PY'
import json
for l in open("gpurun_out/r2_configs20.jsonl"):
    d=json.loads(l); print(d["config"], "%.1f ms"%(d["seconds"]*1e3), "%.3e"%d["path_steps_per_s"], d["launches"], d["timings"])
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_cfg2d.csv python tools/run_configs.py 2 --repeats 1 > /dev/null 2>&1
grep "irc_value_kernel" gpurun_out/r2_launches_cfg2d.csv | awk -F'","' '{print $5, $(NF)}' | cut -c1-120
