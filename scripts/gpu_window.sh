python -m pytest tests/test_gpu_parity.py tests/test_host_wrapper.py tests/test_gpu_edge_cases.py -m gpu -x -q -k "explicit or window or blocked or convergence or dropin or single_camera or ragged or no_obs or eval or determin" > gpurun_out/gputests.log 2>&1; echo rc=$? >> gpurun_out/gputests.log
python bench.py --workload cfg1 --no-cpu-baseline --steps 30 > gpurun_out/bench_cfg1_auto.log 2>&1
python bench.py --workload cfg2 --no-cpu-baseline --steps 30 > gpurun_out/bench_cfg2_auto.log 2>&1
BA_NO_FORK=1 python bench.py --workload cfg1 --no-cpu-baseline --steps 30 >> gpurun_out/bench_cfg1_auto.log 2>&1
BA_NO_FORK=1 python bench.py --workload cfg2 --no-cpu-baseline --steps 30 >> gpurun_out/bench_cfg2_auto.log 2>&1
