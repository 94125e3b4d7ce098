# ncu --set full of the two largest kernels of the windowed LM iteration (cfg 2, REF cost)
BA_NO_LM_GRAPH=1 python profiles/profile_target.py 2 4 > gpurun_out/plain_w.log 2>&1 && \
BA_NO_LM_GRAPH=1 ncu --set full --clock-control none --import-source on -k regex:'k_ldlt2_solve|k_schur_pairs' -s 2 -c 2 -o gpurun_out/r01_prof_window -f python profiles/profile_target.py 2 4 > gpurun_out/ncu_w.log 2>&1
tail -2 gpurun_out/plain_w.log gpurun_out/ncu_w.log
