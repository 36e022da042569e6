python -m pytest tests -x -q -m gpu > gpurun_out/t_all.log 2>&1; tail -3 gpurun_out/t_all.log
python scripts/run_op.py t3ff 1000 4 > gpurun_out/m3.log 2>&1
cat gpurun_out/m3.log
