python scripts/e2e_breakdown.py 1000 > gpurun_out/e2e_breakdown.log 2>&1
tail -40 gpurun_out/e2e_breakdown.log
nproc; free -g | head -2
