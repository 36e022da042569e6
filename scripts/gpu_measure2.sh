python -m pytest tests/test_gpu_multi.py -q -m gpu > gpurun_out/t_multi.log 2>&1; tail -3 gpurun_out/t_multi.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench_r01_2gpu.json 2> gpurun_out/bench_r01_2gpu.err; tail -c 400 gpurun_out/bench_r01_2gpu.err
