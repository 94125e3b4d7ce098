N=${1:-2}
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_cfg5_n${N}_auto.log 2>&1; echo rc=$? >> gpurun_out/bench_cfg5_n${N}_auto.log
