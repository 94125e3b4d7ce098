#!/bin/bash
# round 2 measurement batch (one B200): full GPU test suite, every workload's bench line (with the CPU baseline),
# golden LM traces for parity_vs_n1, tolerance probe, sliding-window schedule with / without the device store,
# ncu launch list of the default bench command and a full capture of the dominant kernels.  Everything lands in
# gpurun_out/ (copied to profiles/ by hand afterwards).
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
O=gpurun_out
nproc > $O/r02_host.txt; nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,power.limit --format=csv >> $O/r02_host.txt
timeout 2400 python -m pytest tests -m gpu -q > $O/r02_gputests.log 2>&1; echo rc=$? >> $O/r02_gputests.log
tail -3 $O/r02_gputests.log
for wl in cfg5 cfg4 cfg3; do
  timeout 900 python bench.py --steps 30 --warmup 0 --workload $wl --no-cpu-baseline --write-golden > $O/r02_golden_$wl.log 2>&1
done
cp tests/golden/bench_trace.json $O/bench_trace.json
for wl in cfg1 cfg2 cfg3 cfg3ref cfg4 cfg5; do
  timeout 1200 python bench.py --steps 20 --warmup 5 --workload $wl > $O/r02_bench_$wl.log 2>&1; echo "$wl rc=$?"
done
timeout 600 python bench.py --steps 20 --warmup 5 --workload cfg5 --solver sparse --no-cpu-baseline > $O/r02_bench_cfg5_pcg.log 2>&1
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > $O/r02_bench_cfg5_reference_arm.log 2>&1
timeout 600 python scripts/chol_tolerance_probe.py > $O/r02_chol_tolerance.txt 2>&1
python - > $O/r02_store_sliding.log 2>&1 <<'PY'
import sys
sys.path.insert(0, '.')
import ba_b200
syn = ba_b200.synthetic
for growing in (False, True):
    for store in (False, True):
        for rep in range(3):
            seq = syn.make_config(2)
            r = ba_b200.hostlib.sliding_sequence(seq, 20, 10, max_num_iterations=10, fixed_iterations=True, device_store=store, growing_maps=growing)
        w = r["windows"]
        print("growing_maps=%d device_store=%d: %d windows, %.1f windows/s | ms per window: " % (growing, store, w, 1e3 * w / r["ms"]["total"])
              + " ".join("%s %.3f" % (k, v / w) for k, v in r["ms"].items()), flush=True)
PY
cat $O/r02_store_sliding.log
CMD="python bench.py --steps 3 --warmup 1 --no-cpu-baseline"
$CMD > $O/r02_ncu_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $O/r02_launches_bench_cfg5.csv $CMD > $O/r02_ncu1.log 2>&1
CMD2="python profiles/profile_target.py 5 3 500 0"
$CMD2 > $O/r02_ncu_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"k_sp_schur|k_spchol_tree|kf_pt_blocks|kf_linearize|kf_schur_pass1|kf_cam_blocks|kf_model_cost|kf_point_inverse" -s 40 -c 24 -o $O/r02_prof_cfg5 $CMD2 > $O/r02_ncu2.log 2>&1
tail -n 2 $O/r02_ncu1.log; tail -n 2 $O/r02_ncu2.log; ls -la $O | head -40; du -sh $O
python scripts/summarize_launches.py $O/r02_launches_bench_cfg5.csv > $O/r02_launches_bench_cfg5_summary.txt 2>&1; head -30 $O/r02_launches_bench_cfg5_summary.txt
for f in $O/r02_bench_*.log; do python - "$f" <<'PY'
import json,sys
for l in open(sys.argv[1]):
    if l.startswith('{'):
        d=json.loads(l); cb=d.get("cpu_baseline") or {}
        print(sys.argv[1].split('/')[-1], "value %.1f e2e %.1f cpu %s final %.12g parity %s" % (d["value"], d["e2e"]["value"], cb.get("value"), d.get("final_cost", 0), d.get("parity_vs_n1")))
PY
done
