python -m pytest tests/test_gpu_parity.py tests/test_host_wrapper.py -m gpu -x -q -k "blocked or explicit or global or window" > gpurun_out/gputests_c2.log 2>&1; echo rc=$? >> gpurun_out/gputests_c2.log
tail -4 gpurun_out/gputests_c2.log
python bench.py --workload cfg3ref --no-cpu-baseline --steps 5 > gpurun_out/c2_cfg3ref.log 2>&1
BA_LEGACY_CHOL=1 python bench.py --workload cfg3ref --no-cpu-baseline --steps 5 > gpurun_out/c2_cfg3ref_legacy.log 2>&1
grep -h '"value"' gpurun_out/c2_*.log | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['config']['workload'][:8], d['value'], d['e2e']['value'], d['final_cost'])
"
ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 1200 --csv --log-file gpurun_out/r01_launches_cfg3ref_v2_warm.csv python bench.py --workload cfg3ref --no-cpu-baseline --steps 2 --warmup 3 > gpurun_out/ncu_ref800b.log 2>&1
