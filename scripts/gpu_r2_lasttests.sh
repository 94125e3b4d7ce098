#!/bin/bash
# the GPU test suite at HEAD (log committed as profiles/r02_gputests.log)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 150 python -m pytest tests -m gpu -q > gpurun_out/r02_gputests.log 2>&1; echo rc=$? >> gpurun_out/r02_gputests.log
tail -n 3 gpurun_out/r02_gputests.log
