# regression check of HEAD: whole single-GPU suite, the default bench command, and the windowed / global REF workloads
t0=$(date +%s)
timeout 420 python -m pytest tests -m gpu -x -q > gpurun_out/head_gputests.log 2>&1; echo rc=$? >> gpurun_out/head_gputests.log
echo "tests $(( $(date +%s) - t0 )) s" > gpurun_out/head_times.log
python bench.py > gpurun_out/head_cfg5.log 2>&1
echo "cfg5 $(( $(date +%s) - t0 )) s" >> gpurun_out/head_times.log
python bench.py --workload cfg1 --no-cpu-baseline --steps 30 > gpurun_out/head_cfg1.log 2>&1
python bench.py --workload cfg2 --no-cpu-baseline --steps 30 > gpurun_out/head_cfg2.log 2>&1
python bench.py --workload cfg3ref --no-cpu-baseline > gpurun_out/head_cfg3ref.log 2>&1
echo "all $(( $(date +%s) - t0 )) s" >> gpurun_out/head_times.log
tail -3 gpurun_out/head_gputests.log; cat gpurun_out/head_times.log
grep -h '"value"' gpurun_out/head_cfg*.log | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['config']['workload'][:8], d['value'], d['e2e']['value'], d['roofline']['kernel'], d['roofline']['frac'])
"
