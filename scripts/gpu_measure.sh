python -m pytest tests -x -q -m gpu -k "explicit or omega or spmv" > gpurun_out/t_expl.log 2>&1; tail -3 gpurun_out/t_expl.log
python scripts/run_op.py explicit 1000 1 > gpurun_out/m5.log 2>&1
cat gpurun_out/m5.log
