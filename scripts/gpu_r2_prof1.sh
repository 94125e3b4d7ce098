#!/bin/bash
# launch list + full ncu capture of the sparse Cholesky kernels and k_sp_schur at cfg 5 (solver 4, 3 LM iterations)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
CMD="python profiles/profile_target.py 5 3 500 4"
$CMD > gpurun_out/r2_prof1_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches_cfg5_chol.csv $CMD > gpurun_out/r2_prof1_ncu1.log 2>&1
$CMD > gpurun_out/r2_prof1_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"k_spchol_factor|k_sp_schur|k_spchol_solve" -s 12 -c 16 -o gpurun_out/r2_prof_spchol $CMD > gpurun_out/r2_prof1_ncu2.log 2>&1
cat gpurun_out/r2_prof1_plain.log; tail -3 gpurun_out/r2_prof1_ncu1.log gpurun_out/r2_prof1_ncu2.log
python scripts/summarize_launches.py gpurun_out/r2_launches_cfg5_chol.csv 2>/dev/null | head -60
