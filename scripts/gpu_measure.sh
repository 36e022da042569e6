python -m pytest tests -q -m gpu > gpurun_out/t_all.log 2>&1; tail -8 gpurun_out/t_all.log
FSGPU_FORCE_GENERIC=1 python -m pytest tests -q -m gpu > gpurun_out/t_all_generic.log 2>&1; tail -4 gpurun_out/t_all_generic.log
ncu --set full --clock-control none --import-source on -k regex:k_t3_stiffness -c 1 -o gpurun_out/t3_v7 -f python scripts/run_op.py t3ff 1000 1 > gpurun_out/ncu_t3.log 2>&1
