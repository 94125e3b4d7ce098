#!/bin/bash
# multi-GPU: parity tests at 2 ranks, then the strong-scaling bench line at N ranks (N = number of GPUs of the box)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/r2_multi_tests_n${N}.log 2>&1; echo rc=$? >> gpurun_out/r2_multi_tests_n${N}.log
tail -5 gpurun_out/r2_multi_tests_n${N}.log
for n in 1 2 4 8; do
  [ $n -gt $N ] && break
  if [ $n -eq 1 ]; then
    timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_bench_cfg5_n${n}.log 2>&1
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2953$n bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/r2_bench_cfg5_n${n}.log 2>&1
  fi
  echo "N=$n rc=$?"
  python - gpurun_out/r2_bench_cfg5_n${n}.log <<'PY'
import json,sys
for l in open(sys.argv[1]):
    if l.startswith('{'):
        d=json.loads(l); print(d["n_gpus"], d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], d["detail"]["solver"][:40], d["final_cost"], d.get("parity_vs_n1"))
        print({k: round(v,3) for k,v in d.get("phase_ms_per_step",{}).items()})
PY
done
