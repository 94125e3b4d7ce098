# programmatic dependent launches in the windowed LM iteration: whole GPU suite, then cfg1 / cfg2 with and without
timeout 420 python -m pytest tests -m gpu -x -q > gpurun_out/pdl_tests.log 2>&1; echo rc=$? >> gpurun_out/pdl_tests.log
tail -4 gpurun_out/pdl_tests.log
for c in cfg1 cfg2; do
  timeout 200 python bench.py --workload $c --no-cpu-baseline --steps 30 > gpurun_out/pdl_$c.log 2>&1
  BA_NO_PDL=1 timeout 200 python bench.py --workload $c --no-cpu-baseline --steps 30 > gpurun_out/nopdl_$c.log 2>&1
done
grep -h '"value"' gpurun_out/pdl_cfg1.log gpurun_out/nopdl_cfg1.log gpurun_out/pdl_cfg2.log gpurun_out/nopdl_cfg2.log | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['config']['workload'][:5], d['value'], d['ms_per_step'], d['final_cost'], d['e2e']['value'], d.get('sliding_sequence',{}).get('cpp_dropin',{}).get('windows_per_s'))
"
