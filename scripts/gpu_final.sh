# end-of-round check: smoke(), whole single-GPU suite, default bench command, windowed / global REF workloads
t0=$(date +%s)
timeout 200 python __graft_entry__.py smoke > gpurun_out/fin_smoke.log 2>&1; echo rc=$? >> gpurun_out/fin_smoke.log
timeout 420 python -m pytest tests -m gpu -x -q > gpurun_out/fin_gputests.log 2>&1; echo rc=$? >> gpurun_out/fin_gputests.log
python bench.py > gpurun_out/fin_cfg5.log 2>&1
python bench.py --workload cfg2 --no-cpu-baseline > gpurun_out/fin_cfg2.log 2>&1
python bench.py --workload cfg3ref --no-cpu-baseline > gpurun_out/fin_cfg3ref.log 2>&1
echo "all $(( $(date +%s) - t0 )) s"
tail -3 gpurun_out/fin_smoke.log; tail -3 gpurun_out/fin_gputests.log
grep -h '"value"' gpurun_out/fin_cfg*.log | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['config']['workload'][:8], d['value'], d['e2e']['value'], d['roofline']['kernel'], d['roofline']['frac'], d.get('sliding_sequence',{}).get('cpp_dropin',{}).get('windows_per_s'))
"
