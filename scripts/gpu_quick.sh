python -m pytest tests/test_gpu_parity.py tests/test_gpu_edge_cases.py -m gpu -x -q > gpurun_out/gputests.log 2>&1; echo rc=$? >> gpurun_out/gputests.log
for c in 5 3 4; do python profiles/profile_target.py $c 10 500 3; done > gpurun_out/plain_sp10.log 2>&1
for c in 5 3 4; do BA_NO_FORK=1 python profiles/profile_target.py $c 10 500 3; done >> gpurun_out/plain_sp10.log 2>&1
