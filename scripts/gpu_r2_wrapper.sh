#!/bin/bash
# host wrapper GPU tests, then bench/window_bench_gpu (the reference's schedule through the compiled windowOptimize) twice + comparison
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 60 python -m pytest tests/test_host_wrapper.py -m gpu -q -x > gpurun_out/r02_wrapper_tests.log 2>&1; echo rc=$? >> gpurun_out/r02_wrapper_tests.log
tail -n 2 gpurun_out/r02_wrapper_tests.log
{
timeout 20 python scripts/dump_sequence.py /tmp/seq.bin --cfg 2
timeout 20 bench/window_bench_gpu /tmp/seq.bin /tmp/a.bin --iterations 10
timeout 20 bench/window_bench_gpu /tmp/seq.bin /tmp/b.bin --iterations 10
timeout 20 python scripts/dump_sequence.py --compare /tmp/a.bin /tmp/b.bin
} > gpurun_out/r02_window_bench_gpu.log 2>&1
cat gpurun_out/r02_window_bench_gpu.log
