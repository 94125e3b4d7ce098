python profiles/profile_target.py 1 10 > gpurun_out/plain_c1.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01_launches_cfg1.csv python profiles/profile_target.py 1 10 > gpurun_out/ncu_c1.log 2>&1
python profiles/profile_target.py 5 1 40 3 > gpurun_out/plain_pt.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'k_pcg_sparse_persistent|k_sp_schur' -c 2 -o gpurun_out/r01_prof_sparse -f python profiles/profile_target.py 5 1 40 3 > gpurun_out/ncu_f.log 2>&1
