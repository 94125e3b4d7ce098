#!/bin/bash
# bench lines + ncu evidence at HEAD after the recursive-halving warp sums (k_sp_schur, camera-block kernels)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python bench.py --steps 20 --warmup 5 --workload cfg5 > $O/r02_bench_cfg5.log 2>&1; echo "cfg5 rc=$?"
for wl in cfg4 cfg3 cfg1; do
  timeout 200 python bench.py --steps 20 --warmup 5 --workload $wl --no-cpu-baseline > $O/r02_bench_${wl}_head.log 2>&1; echo "$wl rc=$?"
done
CMD="python bench.py --steps 3 --warmup 1 --no-cpu-baseline"
timeout 200 $CMD > $O/r02_ncu_plain.log 2>&1 &&
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $O/r02_launches_bench_cfg5.csv $CMD > $O/r02_ncu1.log 2>&1
python scripts/summarize_launches.py $O/r02_launches_bench_cfg5.csv > $O/r02_launches_bench_cfg5_summary.txt 2>&1; head -n 8 $O/r02_launches_bench_cfg5_summary.txt
CMD2="python profiles/profile_target.py 5 3 500 0"
timeout 200 $CMD2 > $O/r02_ncu_plain2.log 2>&1 &&
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"k_sp_schur|kf_cam_blocks" -s 2 -c 4 -f -o $O/r02_prof_cfg5_head $CMD2 > $O/r02_ncu3.log 2>&1
tail -n 2 $O/r02_ncu3.log; rm -f $O/r02_prof_cfg5.ncu-rep; du -sh $O
for f in $O/r02_bench_cfg5.log $O/r02_bench_cfg4_head.log $O/r02_bench_cfg3_head.log $O/r02_bench_cfg1_head.log; do python - "$f" <<'PY'
import json,sys
for l in open(sys.argv[1]):
    if l.startswith('{'):
        d=json.loads(l); cb=d.get("cpu_baseline") or {}
        print(sys.argv[1].split('/')[-1], "value %.1f e2e %.1f cpu %s final %.12g" % (d["value"], d["e2e"]["value"], cb.get("value"), d.get("final_cost", 0)), d["roofline"]["frac"], d["roofline"].get("traffic"))
        print({k: round(v,3) for k,v in d.get("phase_ms_per_step",{}).items()})
PY
done
