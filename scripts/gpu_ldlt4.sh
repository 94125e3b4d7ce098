timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/l4_tests.log 2>&1; echo rc=$? >> gpurun_out/l4_tests.log
tail -3 gpurun_out/l4_tests.log
timeout 200 python bench.py --workload cfg1 --no-cpu-baseline --steps 30 > gpurun_out/l4_cfg1.log 2>&1
timeout 200 python bench.py --workload cfg2 --no-cpu-baseline --steps 30 > gpurun_out/l4_cfg2.log 2>&1
grep -h '"value"' gpurun_out/l4_cfg1.log gpurun_out/l4_cfg2.log | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['config']['workload'][:7], d['value'], d['ms_per_step'], d['final_cost'])
"
