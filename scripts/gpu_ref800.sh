# launch list of one exact LM step of the reference's global optimisation (800 keyframes, REF cost, blocked Cholesky)
python bench.py --workload cfg3ref --no-cpu-baseline --steps 2 --warmup 3 > gpurun_out/ref800_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 1200 --csv --log-file gpurun_out/r01_launches_cfg3ref_warm.csv python bench.py --workload cfg3ref --no-cpu-baseline --steps 2 --warmup 3 > gpurun_out/ncu_ref800.log 2>&1
tail -c 600 gpurun_out/ref800_plain.log
