export PATH=/usr/local/cuda/bin:$PATH
python -m pytest tests -m gpu -x -q -k "t3 or T3 or stiffness or assembl or c4 or c3" > gpurun_out/t_t3.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/t_t3.log
bash scripts/ab.sh "t3ff 1000 6" base t3old t3sl38 t3sl44 base t3old 2>&1 | tee gpurun_out/ab_t3.log
