# launch list of one exact LM step of the reference's global optimisation (800 keyframes, REF cost, blocked Cholesky with the
# look-ahead schedule and k_chol_solve2) and ncu --set full of its kernels, each after the same command exited 0 without ncu
python bench.py --workload cfg3ref --no-cpu-baseline --steps 2 --warmup 3 > gpurun_out/pc_plain.log 2>&1 && {
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 1700 --csv --log-file gpurun_out/r01_launches_cfg3ref.csv python bench.py --workload cfg3ref --no-cpu-baseline --steps 2 --warmup 3 > gpurun_out/pc_ncu_l.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'k_chol_update|k_chol_trsm2|k_chol_potrf2|k_chol_solve2' -s 60 -c 3 -o gpurun_out/r01_prof_chol_a -f python bench.py --workload cfg3ref --no-cpu-baseline --steps 2 --warmup 3 > gpurun_out/pc_ncu_a.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'k_chol_solve2' -c 1 -o gpurun_out/r01_prof_chol_b -f python bench.py --workload cfg3ref --no-cpu-baseline --steps 2 --warmup 3 > gpurun_out/pc_ncu_b.log 2>&1
}
tail -c 400 gpurun_out/pc_plain.log; tail -2 gpurun_out/pc_ncu_a.log gpurun_out/pc_ncu_b.log; ls -la gpurun_out | tail -8
