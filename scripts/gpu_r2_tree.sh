#!/bin/bash
# persistent tree kernel of the sparse Cholesky vs one launch per level: parity tests, then phase times of both
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_host_wrapper.py -m gpu -x -q -k "spchol or lockstep_exact or auto_picks or store" > gpurun_out/r2_tree_tests.log 2>&1; echo rc=$? >> gpurun_out/r2_tree_tests.log
tail -4 gpurun_out/r2_tree_tests.log
for env in "X=1" "BA_SPCHOL_LEVELS=1"; do
  echo "== $env"
  for c in 5 3 4; do env $env timeout 300 python scripts/phase_probe.py $c 10 4 2>&1 | tail -1; done
done
