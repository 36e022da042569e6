set -x
python -m pytest tests -q -m gpu > gpurun_out/t_all.log 2>&1; tail -3 gpurun_out/t_all.log
python bench.py > gpurun_out/bench_r01_final.json 2> gpurun_out/bench_r01_final.err; tail -c 300 gpurun_out/bench_r01_final.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r01_reference.json 2> gpurun_out/bench_r01_reference.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r01.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --expl-steps 5 --power-its 2 > gpurun_out/bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_q4_stiffness -c 1 -o gpurun_out/q4_v8 -f python scripts/run_op.py q4rs 1000 1 > gpurun_out/ncu_q4.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_t3_stiffness -c 1 -o gpurun_out/t3_v9 -f python scripts/run_op.py t3ff 1000 1 > gpurun_out/ncu_t3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_spmv_step -c 1 -o gpurun_out/expl_v2 -f python scripts/run_op.py explicit 1000 1 > gpurun_out/ncu_expl.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_beam_matrix_coop -c 1 -o gpurun_out/beam_v2 -f python scripts/run_op.py beam 69 1 > gpurun_out/ncu_beam.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -5
