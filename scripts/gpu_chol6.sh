# look-ahead schedule 2 (fused diagonal kernel, three streams): parity tests, then cfg3ref A/B against schedule 1
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_host_wrapper.py -m gpu -x -q -k "blocked or cpp_dropin" > gpurun_out/c6_tests.log 2>&1; echo rc=$? >> gpurun_out/c6_tests.log
tail -4 gpurun_out/c6_tests.log
timeout 200 python bench.py --workload cfg3ref --no-cpu-baseline > gpurun_out/c6_new.log 2>&1; echo rc=$?
BA_LOOKAHEAD1=1 timeout 200 python bench.py --workload cfg3ref --no-cpu-baseline > gpurun_out/c6_v1.log 2>&1
grep -h '"value"' gpurun_out/c6_new.log gpurun_out/c6_v1.log | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['value'], d['ms_per_step'], d['final_cost'], d['e2e']['value'])
"
