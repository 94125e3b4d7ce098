# experimental six-lanes-per-block product kernel: parity of the launch-per-step PCG + product hooks, then its time at cfg5
BA_SPMV6=1 timeout 200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_edge_cases.py -m gpu -x -q -k "sparse or product or matvec" > gpurun_out/s6_tests.log 2>&1; echo rc=$? >> gpurun_out/s6_tests.log
tail -3 gpurun_out/s6_tests.log
BA_SPMV6=1 timeout 300 python bench.py --no-cpu-baseline --steps 3 --warmup 3 > gpurun_out/s6_cfg5.log 2>&1; echo rc=$?
grep -h '"value"' gpurun_out/s6_cfg5.log | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['value'], json.dumps(d['roofline_kernels']))
"
