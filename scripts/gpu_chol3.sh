# k_chol_solve2 (flag-in-data blocked substitution): parity tests of the blocked path, then cfg3ref with it and with the
# grid-barrier version (BA_LEGACY_CHOL=3)
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_host_wrapper.py -m gpu -x -q -k "blocked or cpp_dropin" > gpurun_out/c3_tests.log 2>&1; echo rc=$? >> gpurun_out/c3_tests.log
tail -4 gpurun_out/c3_tests.log
timeout 200 python bench.py --workload cfg3ref --no-cpu-baseline > gpurun_out/c3_new.log 2>&1; echo rc=$?
BA_LEGACY_CHOL=3 timeout 200 python bench.py --workload cfg3ref --no-cpu-baseline > gpurun_out/c3_old.log 2>&1
grep -h '"value"' gpurun_out/c3_new.log gpurun_out/c3_old.log | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['value'], d['ms_per_step'], d['final_cost'], d['e2e']['value'])
"
