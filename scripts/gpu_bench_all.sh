python -m pytest tests -m gpu -x -q > gpurun_out/gputests.log 2>&1; echo rc=$? >> gpurun_out/gputests.log
python bench.py > gpurun_out/bench_cfg5_auto.log 2>&1
python bench.py --workload cfg3 --no-cpu-baseline > gpurun_out/bench_cfg3_auto.log 2>&1
python bench.py --workload cfg4 --no-cpu-baseline > gpurun_out/bench_cfg4_auto.log 2>&1
python bench.py --workload cfg1 --no-cpu-baseline > gpurun_out/bench_cfg1_auto.log 2>&1
python bench.py --workload cfg2 --no-cpu-baseline > gpurun_out/bench_cfg2_auto.log 2>&1
