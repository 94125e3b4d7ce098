python -m pytest tests/test_gpu_parity.py tests/test_host_wrapper.py tests/test_gpu_edge_cases.py -m gpu -x -q > gpurun_out/gputests_w5.log 2>&1; echo rc=$? >> gpurun_out/gputests_w5.log
tail -3 gpurun_out/gputests_w5.log
(BA_UPLOAD_PROF=1 python scripts/window_breakdown.py 10 2>&1 | tail -4
echo "--- BA_NO_LM_GRAPH=1"; BA_NO_LM_GRAPH=1 python scripts/window_breakdown.py 10 2>&1 | tail -2
echo "--- BA_HOST_PAIRS=1"; BA_HOST_PAIRS=1 python scripts/window_breakdown.py 10 2>&1 | tail -2
echo "--- 30 iterations"; python scripts/window_breakdown.py 30 2>&1 | tail -1) > gpurun_out/window_breakdown2.log 2>&1
cat gpurun_out/window_breakdown2.log
