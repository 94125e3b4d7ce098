"""Point-sharded multi-GPU solve == single-GPU solve (needs >= 2 GPUs)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _n_gpus():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


@pytest.mark.parametrize("solver,persistent", [(4, 1), (4, 3), (0, 1), (2, 1), (3, 1), (3, 2)],
                         ids=["sparse-cholesky", "sparse-cholesky-distributed", "auto", "implicit", "block-sparse-row-sharded-pcg",
                              "block-sparse-replicated-pcg"])
@pytest.mark.parametrize("cfg", [3, 4])
@pytest.mark.parametrize("world", [2])
def test_sharded_solve_matches_single_gpu(world, cfg, solver, persistent):
    if _n_gpus() < world:
        pytest.skip("needs %d GPUs" % world)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr",
           "127.0.0.1", "--master-port", "29517", os.path.join(ROOT, "tests", "multi_gpu_worker.py"), str(cfg), "0.1" if solver in (0, 4) else "0.05", "6", str(solver), str(persistent)]
    env = dict(os.environ)
    if persistent == 3:
        # small leaves: a deep tree at the reduced test size, so that the subtree-to-rank partition (one subtree group per rank,
        # top part replicated, sparse exchange of S) is really in force -- the worker asserts it
        cmd[-1] = "1"
        cmd.append("distributed")
        env["BA_SPCHOL_LEAF"] = "6"
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "MULTI_GPU_PARITY OK" in out.stdout
