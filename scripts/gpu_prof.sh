# Mid-round profiling pass: step / kernel times of the matrix operators + ncu --set full captures of the Q4 and T3 kernels.
export PATH=/usr/local/cuda/bin:$PATH
python scripts/time_ops.py q4rs t3ff c3 > gpurun_out/time_ops.log 2>&1; cat gpurun_out/time_ops.log | grep -v "^$" | tail -5
ncu --set full --clock-control none --import-source on -k regex:k_q4_stiffness -s 2 -c 1 -o gpurun_out/q4_cur -f python scripts/run_op.py q4rs 1000 4 > gpurun_out/ncu_q4.log 2>&1; echo "ncu q4 rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_t3_stiffness -s 2 -c 1 -o gpurun_out/t3_cur -f python scripts/run_op.py t3ff 1000 4 > gpurun_out/ncu_t3.log 2>&1; echo "ncu t3 rc=$?"
for r in q4_cur t3_cur; do python scripts/ncu_summary.py gpurun_out/$r.ncu-rep > gpurun_out/${r}_summary.txt 2>&1; done
