"""Row N3 of SURVEY.md 8f: batched back-projection / landmark initialisation
(getLocalPoints3D src/Map3D.cpp:76-97, addNewLandmark :44)."""
import numpy as np
import pytest

from helpers import ba_b200, ora


def _case(n, seed, w=640, h=480):
    rng = np.random.default_rng(seed)
    uv = np.stack([rng.uniform(0, w - 1e-3, n), rng.uniform(0, h - 1e-3, n)], 1).astype(np.float32)
    uv[0] = [0.0, 0.0]
    uv[1] = [np.float32(w - 1) + np.float32(0.99), np.float32(h - 1) + np.float32(0.99)]   # trunc -> last pixel
    raw = rng.integers(0, 5000 * 8, size=(h, w)).astype(np.uint16)                        # Kinect raw, 1/5000 m
    raw[rng.random(size=(h, w)) < 0.1] = 0                                                  # holes: depth 0
    img = (raw.astype(np.float32) / np.float32(5000.0)).astype(np.float32)                 # headers/VirtualSensor.h:97-102
    intr = np.array([525.0, 525.0, 319.5, 239.5])
    pose = ba_b200.se3.exp(rng.normal(size=6) * 0.5)
    return uv, img, intr, pose


def test_oracle_backproject_formula():
    """The C oracle against the formula written out in numpy (same operation order => equal bits)."""
    uv, img, intr, pose = _case(500, 0)
    rc, local, world = ora.backproject(uv, img, intr, pose)
    assert rc == 0
    u, v = uv[:, 0].astype(np.float64), uv[:, 1].astype(np.float64)
    z = img[np.trunc(uv[:, 1]).astype(int), np.trunc(uv[:, 0]).astype(int)].astype(np.float64)
    ref = np.stack([z * (u - intr[2]) / intr[0], z * (v - intr[3]) / intr[1], z], 1)
    assert np.array_equal(local, ref)
    for i in range(0, 500, 37):
        assert np.array_equal(world[i], ora.se3_act(pose, local[i]))
    # holes stay at the camera centre (and are later rejected by countConstraints, depth <= 1e-15)
    assert np.all(local[z == 0.0] == 0.0)
    bad = uv.copy()
    bad[3] = [640.0, 10.0]
    assert ora.backproject(bad, img, intr, pose)[0] == -1


@pytest.mark.gpu
@pytest.mark.parametrize("n", [1, 500, 200000])
def test_gpu_backproject_bit_exact(n):
    uv, img, intr, pose = _case(max(n, 2), 11)
    uv = uv[:n] if n >= 2 else uv[:1]
    s = ba_b200.GpuSolver()
    try:
        local, world = s.backproject(uv, img, intr, pose)
        rc, olocal, oworld = ora.backproject(uv, img, intr, pose)
        assert rc == 0
        assert np.array_equal(local, olocal)      # bit-exact: same fp64 operations, no FMA contraction
        assert np.array_equal(world, oworld)
        only_local = s.backproject(uv, img, intr)
        assert np.array_equal(only_local, olocal)
        bad = uv.copy()
        bad[0] = [-1.0, 5.0]
        with pytest.raises(ba_b200.BAError) as e:
            s.backproject(bad, img, intr, pose)
        assert e.value.code == ba_b200.capi.BA_ERR_INVALID
    finally:
        s.close()
