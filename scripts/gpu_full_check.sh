# whole single-GPU suite + the default bench command + cfg 3/4 (regression check after kernel changes)
python -m pytest tests -m gpu -x -q > gpurun_out/gputests_full.log 2>&1; echo rc=$? >> gpurun_out/gputests_full.log
python bench.py > gpurun_out/full_cfg5.log 2>&1
python bench.py --workload cfg3 --no-cpu-baseline > gpurun_out/full_cfg3.log 2>&1
python bench.py --workload cfg4 --no-cpu-baseline > gpurun_out/full_cfg4.log 2>&1
tail -3 gpurun_out/gputests_full.log
grep -h '"value"' gpurun_out/full_cfg*.log | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['config']['workload'][:8], d['value'], d['e2e']['value'], d['roofline']['kernel'], d['roofline']['frac'])
"
