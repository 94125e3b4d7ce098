# windowed explicit path: parity tests, A/B of the three single-CTA solves, warm launch list
python -m pytest tests/test_gpu_parity.py tests/test_host_wrapper.py tests/test_gpu_edge_cases.py -m gpu -x -q > gpurun_out/gputests_w4.log 2>&1; echo rc=$? >> gpurun_out/gputests_w4.log
for c in cfg1 cfg2; do
  python bench.py --workload $c --no-cpu-baseline --steps 30 > gpurun_out/w4_${c}_default.log 2>&1
  BA_NO_FORK=1 python bench.py --workload $c --no-cpu-baseline --steps 30 > gpurun_out/w4_${c}_nofork.log 2>&1
  BA_NO_LM_GRAPH=1 python bench.py --workload $c --no-cpu-baseline --steps 30 > gpurun_out/w4_${c}_nograph.log 2>&1
done
python bench.py --workload cfg3ref --no-cpu-baseline --steps 5 > gpurun_out/w4_cfg3ref.log 2>&1
BA_NO_LM_GRAPH=1 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 400 --csv --log-file gpurun_out/r01_launches_cfg2_v4_warm.csv python bench.py --workload cfg2 --no-cpu-baseline --steps 10 > gpurun_out/ncu_w4b.log 2>&1
BA_NO_LM_GRAPH=1 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 300 --csv --log-file gpurun_out/r01_launches_cfg1_v4_warm.csv python bench.py --workload cfg1 --no-cpu-baseline --steps 10 > gpurun_out/ncu_w4c.log 2>&1
grep -h '"value"' gpurun_out/w4_*.log | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['config']['workload'][:8], d['value'], d['e2e']['value'], d.get('sliding_sequence', {}).get('lm_iterations_per_s'))
"
tail -3 gpurun_out/gputests_w4.log
