python -m pytest tests/test_gpu_parity.py tests/test_host_wrapper.py tests/test_gpu_edge_cases.py -m gpu -x -q > gpurun_out/gputests_w6.log 2>&1; echo rc=$? >> gpurun_out/gputests_w6.log
tail -15 gpurun_out/gputests_w6.log
for c in cfg1 cfg2; do
  python bench.py --workload $c --no-cpu-baseline --steps 30 > gpurun_out/w6_${c}_default.log 2>&1
done
grep -h '"value"' gpurun_out/w6_*.log | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['config']['workload'][:8], d['value'], d['e2e']['value'], d.get('sliding_sequence', {}))
"
python scripts/window_breakdown.py 10 2>&1 | tail -1
