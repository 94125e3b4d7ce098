# after removing look-ahead schedule 1: blocked Cholesky tests (incl. schedule bit-identity) + cfg3ref
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_host_wrapper.py -m gpu -x -q -k "blocked or cpp_dropin" > gpurun_out/c8_tests.log 2>&1; echo rc=$? >> gpurun_out/c8_tests.log
tail -5 gpurun_out/c8_tests.log
timeout 200 python bench.py --workload cfg3ref --no-cpu-baseline > gpurun_out/c8_new.log 2>&1; echo rc=$?
grep -h '"value"' gpurun_out/c8_new.log | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['value'], d['ms_per_step'], d['final_cost'], d['e2e']['value'])
"
