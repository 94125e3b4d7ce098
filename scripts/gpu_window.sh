python -m pytest tests/test_gpu_parity.py tests/test_host_wrapper.py tests/test_gpu_edge_cases.py -m gpu -x -q -k "explicit or window or blocked or convergence or dropin or single_camera or ragged or no_obs" > gpurun_out/gputests.log 2>&1; echo rc=$? >> gpurun_out/gputests.log
python profiles/profile_target.py 1 10 > gpurun_out/plain_c12.log 2>&1
python profiles/profile_target.py 2 10 >> gpurun_out/plain_c12.log 2>&1
