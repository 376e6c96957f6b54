"""CPU test of the Voronoi cell core: tests/host/voronoi_host_harness.cpp builds the __host__ __device__ core of
mdapy_b200/csrc/voronoi_core.cuh for the CPU (test-only) and the result is compared with the golden vectors
(tests/golden/voronoi.npz: the reference's OVITO fixtures and outputs of its own voro++ build).  The GPU runs the
same functions inside k_voronoi (tests/test_gpu_voronoi.py)."""
import ctypes as C
import subprocess
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
GOLD_DIR = ROOT / "tests" / "golden"
GOLD = np.load(GOLD_DIR / "voronoi.npz")
dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int)


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    out = tmp_path_factory.mktemp("voro") / "libvoro_host.so"
    cxx = "/usr/bin/g++" if Path("/usr/bin/g++").exists() else "g++"
    subprocess.run([cxx, "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-o", str(out),
                    str(ROOT / "tests" / "host" / "voronoi_host_harness.cpp")], check=True)
    return C.CDLL(str(out))


def cells(lib, pos, box, origin, boundary, W=40, scale=1.0):
    x, y, z = (np.ascontiguousarray(pos[:, k], np.float64) for k in range(3))
    N = x.shape[0]
    b = np.ascontiguousarray(np.asarray(box, float)[:3].reshape(9))
    o, p = np.ascontiguousarray(origin, np.float64), np.ascontiguousarray(boundary, np.int32)
    vol, nn, rad = np.zeros(N), np.zeros(N, np.int32), np.zeros(N)
    ids, area = np.full((N, W), -1, np.int32), np.zeros((N, W))
    rc = lib.voronoi_host(x.ctypes.data_as(dp), y.ctypes.data_as(dp), z.ctypes.data_as(dp), N, b.ctypes.data_as(dp),
                          o.ctypes.data_as(dp), p.ctypes.data_as(ip), C.c_double(scale), vol.ctypes.data_as(dp),
                          nn.ctypes.data_as(ip), rad.ctypes.data_as(dp), ids.ctypes.data_as(ip), area.ctypes.data_as(dp), W)
    assert rc >= 0, rc
    return vol, nn, rad, ids, area


@pytest.mark.parametrize("name", [str(n) for n in GOLD["fixture_names"]])
def test_upstream_fixture(harness, name):
    d = np.load(GOLD_DIR / f"sa_{name}.npz")
    box = np.asarray(d["box"], float)
    origin = box[3] if box.shape[0] == 4 else np.zeros(3)
    vol, nn, rad, _, _ = cells(harness, d["pos"], box, origin, d["boundary"])       # 3 of the 15 are triclinic
    assert np.array_equal(nn, GOLD[f"{name}__voronoi_coord"])          # perfect lattices included: 12 / 14 / 16 faces
    assert np.allclose(vol, GOLD[f"{name}__voronoi_volume"], atol=1e-6)
    assert np.allclose(rad * 0.5, GOLD[f"{name}__voronoi_cavity_radius"], atol=1e-6)


@pytest.mark.parametrize("scale", [1.0, 0.6, 1.7])
@pytest.mark.parametrize("name", [str(n) for n in GOLD["run_names"]])
def test_reference_run_vectors(harness, name, scale):
    """The cells do not depend on the candidate grid (cell width scaled by 0.6 / 1.7: more shells / bigger shells)."""
    pos, box, bd = GOLD[f"run_{name}__pos"], GOLD[f"run_{name}__box"], GOLD[f"run_{name}__boundary"]
    if np.abs(box - np.diag(np.diag(box))).max() > 1e-10 and not all(bd):
        box, bd = box.copy(), np.ones(3, np.int32)      # mdapy_b200/voronoi.py: open triclinic axes are tripled
        for k in range(3):
            if GOLD[f"run_{name}__boundary"][k] == 0:
                box[k] *= 3
    vol, nn, rad, ids, area = cells(harness, pos, box, np.zeros(3), bd, scale=scale)
    assert np.array_equal(nn, GOLD[f"run_{name}__faces"])
    assert np.allclose(vol, GOLD[f"run_{name}__volume"], rtol=1e-9, atol=0)
    assert np.allclose(rad, GOLD[f"run_{name}__radius"], rtol=1e-9, atol=0)
    # rows as sets: neighbour ids (walls = -1) and face areas against the unfiltered reference rows
    rv, ra = GOLD[f"run_{name}__none_verlet"], GOLD[f"run_{name}__none_area"]
    M = rv.shape[1]
    ids = np.where(ids < 0, -1, ids)[:, :M]
    area = np.where(ids < 0, 0.0, area[:, :M])          # the reference zeroes wall faces (voronoi.cpp:421-426)
    key = np.where(ids < 0, np.iinfo(np.int32).max, ids)
    order = np.lexsort((area, key), axis=1)
    assert np.array_equal(np.take_along_axis(ids, order, axis=1), rv)
    assert np.allclose(np.take_along_axis(area, order, axis=1), ra, rtol=1e-7, atol=1e-9)


def test_fuzz_against_compiled_reference(harness):
    """Random boxes (3..30 per axis, any boundary mix), 2..400 atoms: uniform, clustered, lattice + noise of
    1e-8..1e-1, far periodic images -- face counts equal, volume and radius to 1e-9 against the reference's voro++."""
    from oracle import ref

    if not ref.available():
        pytest.skip("needs oracle/_ref (the reference compiled in the dev container)")
    rng = np.random.default_rng(2024)
    for trial in range(80):
        N = int(rng.integers(2, 400))
        L = rng.uniform(3, 30, 3)
        bd = rng.integers(0, 2, 3).astype(np.int32)
        mode = trial % 4
        if mode == 0:
            pos = rng.random((N, 3)) * L
        elif mode == 1:
            c = rng.random((max(1, N // 20), 3)) * L
            pos = c[rng.integers(0, len(c), N)] + rng.normal(0, 0.4, (N, 3))
            pos = np.where(bd == 1, pos % L, np.clip(pos, 1e-6, L - 1e-6))
        elif mode == 2:
            n = int(round(N ** (1 / 3))) + 1
            g = np.stack(np.meshgrid(*[np.arange(n)] * 3, indexing="ij"), -1).reshape(-1, 3)[:N]
            pos = (g + 0.5) * (L / n) + rng.normal(0, 10 ** rng.uniform(-8, -1), (len(g), 3))
            pos = np.where(bd == 1, pos % L, np.clip(pos, 1e-6, L - 1e-6))
        else:
            pos = rng.random((N, 3)) * L
            pos = np.where(bd == 1, pos + rng.integers(-2, 3, (N, 3)) * L, pos)
        box = np.diag(L)
        x, y, z = (np.ascontiguousarray(pos[:, k]) for k in range(3))
        rvol, rnn, rrad = ref.voronoi_volume(x, y, z, box, np.zeros(3), bd)
        vol, nn, rad, _, _ = cells(harness, pos, box, np.zeros(3), bd, W=64)
        m = rvol > 0
        assert np.array_equal(nn, rnn), (trial, mode)
        assert np.allclose(vol[m], rvol[m], rtol=1e-9, atol=0), (trial, mode)
        assert np.allclose(rad[m], rrad[m], rtol=1e-9, atol=0), (trial, mode)


def test_fuzz_triclinic_against_compiled_reference(harness):
    """Sheared boxes (tilt factors up to +-0.5), periodic and open axes (open axes tripled like the reference's
    Python side): uniform points and full lattices with 1e-8..1e-2 noise against get_voronoi_volume_number_radius_tri."""
    from oracle import ref

    if not ref.available():
        pytest.skip("needs oracle/_ref (the reference compiled in the dev container)")
    rng = np.random.default_rng(77)
    for trial in range(60):
        L = rng.uniform(6, 25, 3)
        F = np.array([[1, 0, 0], [rng.uniform(-0.5, 0.5), 1, 0], [rng.uniform(-0.5, 0.5), rng.uniform(-0.5, 0.5), 1]])
        box = np.diag(L) @ F
        bd = np.array([1, 1, 1], np.int32) if trial % 3 else rng.integers(0, 2, 3).astype(np.int32)
        if trial % 2:
            n = int(rng.integers(3, 7))
            g = np.stack(np.meshgrid(*[np.arange(n)] * 3, indexing="ij"), -1).reshape(-1, 3)
            frac = ((g + 0.5) / n + rng.normal(0, 10 ** rng.uniform(-8, -2), g.shape)) % 1.0
        else:
            frac = rng.random((int(rng.integers(20, 400)), 3))
        pos = frac @ box
        x, y, z = (np.ascontiguousarray(pos[:, k]) for k in range(3))
        rvol, rnn, rrad = ref.voronoi_volume_tri(x, y, z, box, np.zeros(3), bd)
        cell = box.copy()
        for k in range(3):
            if bd[k] == 0:
                cell[k] *= 3
        vol, nn, rad, _, _ = cells(harness, pos, cell, np.zeros(3), [1, 1, 1], W=64)
        assert np.array_equal(nn, rnn), trial
        assert np.allclose(vol, rvol, rtol=1e-9, atol=0) and np.allclose(rad, rrad, rtol=1e-9, atol=0), trial
