# iteration zero with PDL + the wrapper without the counting pre-pass: whole GPU suite, cfg1 / cfg2 bench
timeout 420 python -m pytest tests -m gpu -x -q > gpurun_out/w7_tests.log 2>&1; echo rc=$? >> gpurun_out/w7_tests.log
tail -3 gpurun_out/w7_tests.log
timeout 200 python bench.py --workload cfg1 --no-cpu-baseline --steps 30 > gpurun_out/w7_cfg1.log 2>&1
timeout 200 python bench.py --workload cfg2 --no-cpu-baseline > gpurun_out/w7_cfg2.log 2>&1
grep -h '"value"' gpurun_out/w7_cfg1.log gpurun_out/w7_cfg2.log | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['config']['workload'][:7], d['value'], d['ms_per_step'], d['final_cost'], d['e2e']['value'], json.dumps(d.get('sliding_sequence',{}).get('cpp_dropin',{}).get('ms_per_window')), d.get('sliding_sequence',{}).get('cpp_dropin',{}).get('windows_per_s'))
"
