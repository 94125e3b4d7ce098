#!/bin/bash
# bench/window_bench_gpu: the reference's schedule (src/main.cpp:161-182) through the compiled windowOptimize, warm second pass reported
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
{
timeout 20 python scripts/dump_sequence.py /tmp/seq.bin --cfg 2
timeout 20 bench/window_bench_gpu /tmp/seq.bin /tmp/a.bin --iterations 10
} > gpurun_out/r02_window_bench_gpu.log 2>&1
cat gpurun_out/r02_window_bench_gpu.log
