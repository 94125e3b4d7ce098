# 2-GPU sanity after the stream-priority / PDL changes: sharded-vs-single parity (cfg 3 variants) and the cfg5 bench at N = 2
timeout 300 python -m pytest tests/test_gpu_multi.py -m gpu -x -q -k "2-3-" > gpurun_out/m2_tests.log 2>&1; echo rc=$? >> gpurun_out/m2_tests.log
tail -4 gpurun_out/m2_tests.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/m2_bench.log 2>&1; echo rc=$?
grep -h '"value"' gpurun_out/m2_bench.log | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['n_gpus'], d['value'], d['ms_per_step'], d['final_cost'], d['e2e']['value'])
"
