python profiles/profile_target.py 5 1 40 3 > gpurun_out/plain_pt.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'k_pcg_sparse_persistent|k_sp_schur' -c 2 -o gpurun_out/r01_prof_sparse -f python profiles/profile_target.py 5 1 40 3 > gpurun_out/ncu_f.log 2>&1
python profiles/profile_target.py 5 1 12 2 3 > gpurun_out/plain_pt2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'kt_schur_fused' -s 3 -c 1 -o gpurun_out/r01_prof_tiled -f python profiles/profile_target.py 5 1 12 2 3 > gpurun_out/ncu_f2.log 2>&1
python profiles/profile_target.py 5 1 12 2 2 > gpurun_out/plain_pt3.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'kf_schur_pass' -s 6 -c 2 -o gpurun_out/r01_prof_fact -f python profiles/profile_target.py 5 1 12 2 2 > gpurun_out/ncu_f3.log 2>&1
