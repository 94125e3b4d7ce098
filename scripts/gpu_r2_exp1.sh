#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
for env in "X=1" "BA_SP_SCHUR_2CTA=1" "BA_SP_SCHUR_ROWS=1"; do
  echo "== $env"; env $env python scripts/phase_probe.py 5 10 4 2>&1 | tail -1
  env $env python scripts/phase_probe.py 3 10 4 2>&1 | tail -1
done
