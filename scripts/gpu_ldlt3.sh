# ldlt2_step with split arrive / wait (mbarrier): whole GPU suite, then cfg1 / cfg2 / cfg3ref
timeout 420 python -m pytest tests -m gpu -x -q > gpurun_out/l3_tests.log 2>&1; echo rc=$? >> gpurun_out/l3_tests.log
tail -4 gpurun_out/l3_tests.log
timeout 200 python bench.py --workload cfg1 --no-cpu-baseline --steps 30 > gpurun_out/l3_cfg1.log 2>&1
timeout 200 python bench.py --workload cfg2 --no-cpu-baseline --steps 30 > gpurun_out/l3_cfg2.log 2>&1
timeout 200 python bench.py --workload cfg3ref --no-cpu-baseline > gpurun_out/l3_cfg3ref.log 2>&1
grep -h '"value"' gpurun_out/l3_cfg1.log gpurun_out/l3_cfg2.log gpurun_out/l3_cfg3ref.log | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['config']['workload'][:7], d['value'], d['ms_per_step'], d['final_cost'], d['e2e']['value'], d.get('sliding_sequence',{}).get('cpp_dropin',{}).get('windows_per_s'))
"
