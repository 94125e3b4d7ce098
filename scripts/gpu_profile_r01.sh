set -x
python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/b.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r01_launches_bench_cfg5.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_l.log 2>&1
python profiles/profile_target.py 5 1 40 3 > gpurun_out/plain_pt.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'k_pcg_sparse_persistent|k_sp_schur|kf_linearize' -c 4 -o gpurun_out/r01_prof_sparse -f python profiles/profile_target.py 5 1 40 3 > gpurun_out/ncu_f.log 2>&1
