python -m pytest tests -x -q -m gpu -k "beam or gyro or distrib" > gpurun_out/t_beam.log 2>&1; tail -3 gpurun_out/t_beam.log
python scripts/run_op.py beam 69 3 > gpurun_out/m4.log 2>&1
cat gpurun_out/m4.log
