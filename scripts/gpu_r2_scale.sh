#!/bin/bash
# strong-scaling bench lines of cfg5 at N = 8, 4, 2 ranks of one box (N = 1: scripts/gpu_r2_final.sh), phase breakdown per N
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l)
for n in ${SCALE_NS:-8 4 2}; do
  [ $n -gt $NG ] && continue
  BA_SPCHOL_DEBUG=1 timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2955$n bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/r02_scale_cfg5_n${n}.log 2>&1
  echo "N=$n rc=$?"
  grep "^\[spchol\] rank 0" gpurun_out/r02_scale_cfg5_n${n}.log | head -n 1
  python - gpurun_out/r02_scale_cfg5_n${n}.log <<'PY'
import json,sys
for l in open(sys.argv[1]):
    if l.startswith('{'):
        d=json.loads(l); print(d["n_gpus"], round(d["value"],1), round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1), d["final_cost"], d.get("parity_vs_n1"))
        print({k: round(v,3) for k,v in d.get("phase_ms_per_step",{}).items()})
PY
done
