python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "fetch_narrow" > gpurun_out/t_fetch.log 2>&1; tail -15 gpurun_out/t_fetch.log
python scripts/e2e_breakdown.py 1000 > gpurun_out/e2e_breakdown.log 2>&1
grep -A12 "rep 2" gpurun_out/e2e_breakdown.log; grep e2e_step gpurun_out/e2e_breakdown.log
