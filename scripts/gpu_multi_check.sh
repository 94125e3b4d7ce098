timeout 400 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/gputests_multi.log 2>&1; echo rc=$? >> gpurun_out/gputests_multi.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_cfg5_n2_auto.log 2>&1; echo rc=$? >> gpurun_out/bench_cfg5_n2_auto.log
