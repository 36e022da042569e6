# Round-end GPU pass (one B200): parity suite, smoke, default bench, reference arm, ncu launch list of the bench command,
# ncu --set full captures of the dominant kernels, FP64 peak record, Julia smoke when Julia exists.  Outputs under gpurun_out/.
export PATH=/usr/local/cuda/bin:$PATH
python -m pytest tests -m gpu -x -q > gpurun_out/t_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/t_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
python bench.py > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err; echo "bench rc=$?"; tail -c 300 gpurun_out/bench_1gpu.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "ref rc=$?"
python scripts/fp64_peak.py > gpurun_out/fp64_peak.json 2> gpurun_out/fp64_peak.err; echo "fp64 peak rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --e2e-steps 1 --expl-steps 20 > gpurun_out/bench_under_ncu.log 2>&1; echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_q4_stiffness -s 2 -c 1 -o gpurun_out/q4_final -f python scripts/run_op.py q4rs 1000 4 > gpurun_out/ncu_q4.log 2>&1; echo "ncu q4 rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_t3_stiffness -s 2 -c 1 -o gpurun_out/t3_final -f python scripts/run_op.py t3ff 1000 4 > gpurun_out/ncu_t3.log 2>&1; echo "ncu t3 rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_spmv_step -s 5 -c 1 -o gpurun_out/expl_final -f python scripts/run_op.py explicit 1000 2 > gpurun_out/ncu_expl.log 2>&1; echo "ncu expl rc=$?"
for r in q4_final t3_final expl_final; do python scripts/ncu_summary.py gpurun_out/$r.ncu-rep > gpurun_out/${r}_summary.txt 2>&1; done
if command -v julia >/dev/null 2>&1; then LIBFSGPU=$PWD/finetoolsflexstructures.jl_b200/libfsgpu.so julia finetoolsflexstructures.jl_b200/julia/smoke.jl > gpurun_out/julia_smoke.log 2>&1; echo "julia smoke rc=$?"; else echo "julia: not installed on this box" | tee gpurun_out/julia_smoke.log; fi
ls -la gpurun_out/
