#!/bin/bash
# closing measurement batch of round 2 at HEAD (one B200): smoke, full GPU test suite, bench lines of the workloads the last kernel
# changes touch (cfg 5 / 4 / 3) + the CPU arm on 1 and all threads, ncu launch list of the default bench command and one full capture
# of two LM iterations' dominant kernels.  Every command under its own timeout.
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python __graft_entry__.py smoke > $O/r02_smoke.log 2>&1; echo "smoke rc=$?"; tail -n 3 $O/r02_smoke.log
timeout 900 python -m pytest tests -m gpu -q > $O/r02_gputests.log 2>&1; echo rc=$? >> $O/r02_gputests.log
tail -n 3 $O/r02_gputests.log
for wl in cfg5 cfg4 cfg3; do
  timeout 400 python bench.py --steps 20 --warmup 5 --workload $wl > $O/r02_bench_$wl.log 2>&1; echo "$wl rc=$?"
done
timeout 400 python scripts/cpu_baseline_threads.py > $O/r02_cpu_baseline_threads.jsonl 2>&1; echo "cpu rc=$?"
CMD="python bench.py --steps 3 --warmup 1 --no-cpu-baseline"
timeout 300 $CMD > $O/r02_ncu_plain.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $O/r02_launches_bench_cfg5.csv $CMD > $O/r02_ncu1.log 2>&1
python scripts/summarize_launches.py $O/r02_launches_bench_cfg5.csv > $O/r02_launches_bench_cfg5_summary.txt 2>&1; head -n 16 $O/r02_launches_bench_cfg5_summary.txt
CMD2="python profiles/profile_target.py 5 3 500 0"
timeout 300 $CMD2 > $O/r02_ncu_plain2.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_sp_schur|k_spchol_tree|kf_pt_blocks|kf_linearize|kf_schur_pass1|kf_cam_blocks|kf_model_cost|kf_point_inverse|kf_schur_pass2|k_cost" -s 4 -c 22 -f -o $O/r02_prof_cfg5 $CMD2 > $O/r02_ncu2.log 2>&1
tail -n 2 $O/r02_ncu2.log; du -sh $O
for f in $O/r02_bench_cfg5.log $O/r02_bench_cfg4.log $O/r02_bench_cfg3.log; do python - "$f" <<'PY'
import json,sys
for l in open(sys.argv[1]):
    if l.startswith('{'):
        d=json.loads(l); cb=d.get("cpu_baseline") or {}
        print(sys.argv[1].split('/')[-1], "value %.1f e2e %.1f cpu %s final %.12g parity %s" % (d["value"], d["e2e"]["value"], cb.get("value"), d.get("final_cost", 0), d.get("parity_vs_n1")))
        print({k: round(v,3) for k,v in d.get("phase_ms_per_step",{}).items()})
PY
done
cat $O/r02_cpu_baseline_threads.jsonl | cut -c1-200
