timeout 600 ncu --set full --clock-control none --import-source on -k regex:irc_value_kernel -s 1 -c 1 -f -o gpurun_out/r2_value_cfg2 python tools/run_configs.py 2 --repeats 1 > gpurun_out/r2_ncu_value.log 2>&1
ls -la gpurun_out/r2_value_cfg2.ncu-rep
