# windowed explicit path: parity tests, then A/B of the LM-iteration graph and the L D L^T solve (cfg 1-2)
python -m pytest tests/test_gpu_parity.py tests/test_host_wrapper.py tests/test_gpu_edge_cases.py -m gpu -x -q > gpurun_out/gputests_w2.log 2>&1; echo rc=$? >> gpurun_out/gputests_w2.log
for c in cfg1 cfg2; do
  python bench.py --workload $c --no-cpu-baseline --steps 30 > gpurun_out/w2_${c}_default.log 2>&1
  BA_NO_LM_GRAPH=1 python bench.py --workload $c --no-cpu-baseline --steps 30 > gpurun_out/w2_${c}_nograph.log 2>&1
  BA_NO_LM_GRAPH=1 BA_LEGACY_CHOL=1 python bench.py --workload $c --no-cpu-baseline --steps 30 > gpurun_out/w2_${c}_nograph_legacychol.log 2>&1
done
python bench.py --workload cfg3ref --no-cpu-baseline --steps 5 > gpurun_out/w2_cfg3ref.log 2>&1
BA_NO_LM_GRAPH=1 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01_launches_cfg2_v2.csv python bench.py --workload cfg2 --no-cpu-baseline --steps 10 > gpurun_out/ncu_w2.log 2>&1
BA_NO_LM_GRAPH=1 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 400 --csv --log-file gpurun_out/r01_launches_cfg2_v2_warm.csv python bench.py --workload cfg2 --no-cpu-baseline --steps 10 > gpurun_out/ncu_w2b.log 2>&1
grep -h '"value"' gpurun_out/w2_*.log | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['config']['workload'][:8], d['value'], d['e2e']['value'], d.get('sliding_sequence', {}).get('lm_iterations_per_s'))
"
tail -3 gpurun_out/gputests_w2.log
