BA_PCG_PROF=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/multi_gpu_worker.py 5 1.0 2 3 1 > gpurun_out/dist_prof.log 2>&1; echo rc=$? >> gpurun_out/dist_prof.log
timeout 400 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/gputests_multi.log 2>&1; echo rc=$? >> gpurun_out/gputests_multi.log
bash scripts/gpu_multi_bench.sh 2
