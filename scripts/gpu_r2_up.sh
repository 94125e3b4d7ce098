cd "$GRAFT_REPO_ROOT"
BA_UPLOAD_PROF=1 python - <<'PY' 2>&1 | tail -6
import sys, time
sys.path.insert(0, '.')
import ba_b200, torch, numpy as np
p = ba_b200.synthetic.make_config(5)
for f in ("pose7", "pt3", "cam_idx", "pt_idx", "uv2"):
    setattr(p, f, torch.from_numpy(np.ascontiguousarray(getattr(p, f))).pin_memory().numpy())
s = ba_b200.GpuSolver(max_num_iterations=2, use_depth_prior=0, optimize_intrinsics=0)
for rep in range(3):
    t0 = time.time(); s.upload(p); t1 = time.time(); s.solve(); t2 = time.time(); s.download(); t3 = time.time()
    print("upload %.2f ms, solve(2 it) %.2f, download %.2f, symbolic %.2f ms" % ((t1-t0)*1e3, (t2-t1)*1e3, (t3-t2)*1e3, s.spchol_info()["symbolic_us"]/1e3), flush=True)
PY
