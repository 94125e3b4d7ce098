# (a) the compiled drop-in over a sliding sequence: test + cfg 2 bench line with the C++ phase breakdown
# (b) launch lists (windowed LM iteration without graph replay, global REF BA) and ncu --set full of the windowed /
#     blocked-Cholesky kernels, each after the same command has exited 0 without ncu
timeout 300 python -m pytest tests/test_host_wrapper.py -m gpu -x -q > gpurun_out/fa_hosttests.log 2>&1; echo rc=$? >> gpurun_out/fa_hosttests.log
python bench.py --workload cfg2 --no-cpu-baseline > gpurun_out/fa_cfg2.log 2>&1
BA_NO_LM_GRAPH=1 python profiles/profile_target.py 2 4 > gpurun_out/fa_plain_w.log 2>&1 && {
BA_NO_LM_GRAPH=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01_launches_cfg2_window.csv python profiles/profile_target.py 2 4 > gpurun_out/fa_ncu_wl.log 2>&1
BA_NO_LM_GRAPH=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:'k_ldlt2_solve|k_schur_pairs|k_obs_W' -s 3 -c 3 -o gpurun_out/r01_prof_window -f python profiles/profile_target.py 2 4 > gpurun_out/fa_ncu_w.log 2>&1
}
python bench.py --workload cfg3ref --no-cpu-baseline --steps 2 --warmup 3 > gpurun_out/fa_ref800_plain.log 2>&1 && {
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r01_launches_cfg3ref.csv python bench.py --workload cfg3ref --no-cpu-baseline --steps 2 --warmup 3 > gpurun_out/fa_ncu_ref800.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'k_chol_update|k_chol_trsm2|k_chol_potrf2|k_chol_solve' -s 40 -c 4 -o gpurun_out/r01_prof_chol -f python bench.py --workload cfg3ref --no-cpu-baseline --steps 2 --warmup 3 > gpurun_out/fa_ncu_chol.log 2>&1
}
tail -3 gpurun_out/fa_hosttests.log; tail -c 1500 gpurun_out/fa_cfg2.log; tail -2 gpurun_out/fa_plain_w.log gpurun_out/fa_ncu_w.log gpurun_out/fa_ncu_chol.log
ls -la gpurun_out
