set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2f_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2f_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2f_smoke.log 2>&1
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2f_bench_1gpu.log 2> gpurun_out/r2f_bench_1gpu.err
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2f_bench_reference.log 2> gpurun_out/r2f_bench_reference.err
timeout 1200 python tools/run_configs.py 1 2 2o 3 3g 4 5 5g > gpurun_out/r2f_configs.jsonl 2> gpurun_out/r2f_configs.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:irc_value_kernel -s 1 -c 1 -f -o gpurun_out/r2f_value_cfg2 python tools/run_configs.py 2 --repeats 1 > /dev/null 2>&1
grep -n "passed\|failed" gpurun_out/r2f_pytest.log; cat gpurun_out/r2f_smoke.log | tail -1
python - <<'PY'
import json
for l in open("gpurun_out/r2f_configs.jsonl"):
    d=json.loads(l); print(d["config"], "%.1f ms"%(d["seconds"]*1e3), "%.3e"%d["path_steps_per_s"], d["launches"], d["timings"])
for f in ("gpurun_out/r2f_bench_1gpu.log","gpurun_out/r2f_bench_reference.log"):
    d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, "%.4e"%d["value"], d["ms_per_step"], "e2e %.4e"%d["e2e"]["value"])
PY
