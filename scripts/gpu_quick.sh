python -m pytest tests/test_gpu_parity.py tests/test_gpu_edge_cases.py -m gpu -x -q -k "sparse or ragged or auto or no_obs" > gpurun_out/gputests.log 2>&1; echo rc=$? >> gpurun_out/gputests.log
python profiles/profile_target.py 5 10 500 3 > gpurun_out/plain_sp10.log 2>&1
python profiles/profile_target.py 3 10 500 3 >> gpurun_out/plain_sp10.log 2>&1
python profiles/profile_target.py 4 10 500 3 >> gpurun_out/plain_sp10.log 2>&1
