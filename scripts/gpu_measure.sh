python -m pytest tests -q -m gpu -x > gpurun_out/t_all.log 2>&1; tail -4 gpurun_out/t_all.log
FSGPU_FORCE_GENERIC=1 python -m pytest tests -q -m gpu > gpurun_out/t_all_generic.log 2>&1; tail -4 gpurun_out/t_all_generic.log
python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -3 gpurun_out/smoke.log
