"""GPU tests through the public API (System.cal_*), written like the reference's own fixture-driven
tests: golden labels of tests/golden/sa_*.npz (upstream OVITO/freud vectors), known answers for perfect
crystals, neighbour-list cache semantics (reference tests/test_system.py:188-215, 329-353), and the
brute-force neighbour oracle of tests/test_neighbor_cutoff.py."""
import glob
from pathlib import Path

import numpy as np
import pytest

import helpers as H

pytestmark = pytest.mark.gpu
GOLD = Path(__file__).resolve().parent / "golden"
SA = sorted(glob.glob(str(GOLD / "sa_*.npz")))


def _system(d):
    import mdapy_b200 as mp

    return mp.System(pos=d["pos"], box=mp.Box(d["box"], boundary=list(d["boundary"])))


@pytest.mark.parametrize("path", SA, ids=[Path(p).stem[3:] for p in SA])
def test_golden_fixture(path):
    d = np.load(path)
    if "cna" in d.files:
        s = _system(d)
        s.cal_common_neighbor_analysis(float(d["cna_cutoff"]))
        assert np.array_equal(np.asarray(s.data["cna"]), d["cna"])
    s = _system(d)
    s.cal_common_neighbor_analysis()
    assert np.array_equal(np.asarray(s.data["cna"]), d["ref_acna"])
    s = _system(d)
    s.cal_ackland_jones_analysis()
    assert np.array_equal(np.asarray(s.data["aja"]), d["aja"])
    if "csp" in d.files:
        s = _system(d)
        s.cal_centro_symmetry_parameter(int(d["csp_num_neighbors"]))
        assert np.allclose(np.asarray(s.data["csp"]), d["csp"], atol=1e-6, rtol=1e-6)
    if "q6" in d.files:
        rc = float(d["ql_cutoff"])
        s = _system(d)
        s.cal_steinhardt_bond_orientation([4, 6], rc=rc)
        assert np.allclose(np.asarray(s.data["ql4"]), d["q4"], atol=1e-6, rtol=1e-6)
        assert np.allclose(np.asarray(s.data["ql6"]), d["q6"], atol=1e-6, rtol=1e-6)
        s = _system(d)
        s.cal_steinhardt_bond_orientation([4, 6], rc=rc, average=True)
        assert np.allclose(np.asarray(s.data["ql4"]), d["q4_avg"], atol=1e-6, rtol=1e-6)
        assert np.allclose(np.asarray(s.data["ql6"]), d["q6_avg"], atol=1e-6, rtol=1e-6)
        s = _system(d)
        sbo = s.cal_steinhardt_bond_orientation([4, 6, 8], rc=rc, wl=True, wlhat=True)
        assert np.allclose(sbo.qnarray[: s.N], d["ref_q468_wl_wlhat"], atol=1e-12, rtol=1e-9)


def test_golden_rdf():
    import mdapy_b200 as mp

    d = np.load(GOLD / "rdf_alcrni.npz")
    s = mp.System(data={"x": d["pos"][:, 0], "y": d["pos"][:, 1], "z": d["pos"][:, 2], "element": d["element"]},
                  box=mp.Box(d["box"]))
    rdf = s.cal_radial_distribution_function(float(d["cutoff"]), int(d["nbins"]))
    el = [str(e) for e in d["elements"]]
    for i in range(len(el)):
        for j in range(i, len(el)):
            key = (el[i], el[j]) if (el[i], el[j]) in rdf.g_partial else (el[j], el[i])
            assert np.allclose(rdf.g_partial[key], d["g"][i, j], atol=1e-6), key
    s2 = mp.System(data={"x": d["pos"][:, 0], "y": d["pos"][:, 1], "z": d["pos"][:, 2], "element": d["element"]},
                   box=mp.Box(d["box"]))
    r2 = s2.cal_radial_distribution_function(float(d["cutoff"]), int(d["nbins"]), streaming=True)
    assert np.allclose(r2.g_total, rdf.g_total, atol=1e-9)


def test_known_answers_perfect_crystals():
    import mdapy_b200 as mp

    fcc = mp.build_crystal("Cu", "fcc", 3.615, nx=6, ny=6, nz=6)
    fcc.cal_common_neighbor_analysis(3.615 * 0.8536)
    assert np.all(np.asarray(fcc.data["cna"]) == 1)
    fcc.cal_common_neighbor_analysis()
    assert np.all(np.asarray(fcc.data["cna"]) == 1)
    fcc.cal_ackland_jones_analysis()
    assert np.all(np.asarray(fcc.data["aja"]) == 1)
    fcc.cal_centro_symmetry_parameter(12)
    assert np.allclose(np.asarray(fcc.data["csp"]), 0.0, atol=1e-10)
    fcc.cal_steinhardt_bond_orientation([4, 6], nnn=12)
    assert np.allclose(np.asarray(fcc.data["ql4"]), 0.190941, atol=1e-5)
    assert np.allclose(np.asarray(fcc.data["ql6"]), 0.574524, atol=1e-5)
    bcc = mp.build_crystal("Fe", "bcc", 2.8665, nx=7, ny=7, nz=7)
    bcc.cal_common_neighbor_analysis()
    assert np.all(np.asarray(bcc.data["cna"]) == 3)
    bcc.cal_ackland_jones_analysis()
    assert np.all(np.asarray(bcc.data["aja"]) == 3)


def _bf_neighbors(pos, box, rc):
    """O(N^2) minimum-image oracle for an orthogonal fully periodic box (test_neighbor_cutoff.py:23-37)."""
    L = np.diag(box)
    out = []
    for i in range(pos.shape[0]):
        d = pos - pos[i]
        d -= L * np.round(d / L)
        r = np.sqrt((d * d).sum(1))
        idx = np.nonzero((r <= rc + 1e-9) & (np.arange(pos.shape[0]) != i))[0]
        out.append((set(idx.tolist()), np.sort(r[idx])))
    return out


@pytest.mark.parametrize("max_neigh", [None, 150])
def test_neighbor_class_vs_brute_force(max_neigh):
    import mdapy_b200 as mp

    pos, box = H.random_gas(600, 16.0, 3)
    nb = mp.Neighbor(3.5, mp.Box(box), {"x": pos[:, 0], "y": pos[:, 1], "z": pos[:, 2]}, max_neigh)
    nb.compute()
    bf = _bf_neighbors(pos, box, 3.5)
    for i, (ids, dist) in enumerate(bf):
        k = nb.neighbor_number[i]
        assert set(nb.verlet_list[i, :k].tolist()) == ids
        assert np.allclose(np.sort(nb.distance_list[i, :k]), dist, atol=1e-6)
        assert np.all(nb.verlet_list[i, k:] == -1) and np.all(nb.distance_list[i, k:] == 4.5)


def test_neighbor_validation_and_small_box():
    import mdapy_b200 as mp

    pos, box = H.fcc(3.615, 4)
    data = {"x": pos[:, 0], "y": pos[:, 1], "z": pos[:, 2]}
    with pytest.raises(AssertionError):
        mp.Neighbor(-1.0, mp.Box(box), data)
    with pytest.raises(AssertionError):
        mp.Neighbor(3.0, mp.Box(box), data, max_neigh=0)
    with pytest.raises(AssertionError):
        mp.Neighbor(3.0, mp.Box(box), {"x": pos[:, 0]})
    with pytest.raises(ValueError, match="max_neigh=5 is too small"):
        mp.Neighbor(3.0, mp.Box(box), data, max_neigh=5).compute()
    nb = mp.Neighbor(3.0, mp.Box(box), data, max_neigh=12)   # exactly enough
    nb.compute()
    assert np.all(nb.neighbor_number == 12)
    # box thinner than 2*rc is replicated, original atoms are rows 0..N-1 (neighbor.py:94-101)
    p1, b1 = H.fcc(3.615, 1)
    nb = mp.Neighbor(3.0, mp.Box(b1), {"x": p1[:, 0], "y": p1[:, 1], "z": p1[:, 2]})
    nb.compute()
    assert hasattr(nb, "_enlarge_data") and nb._enlarge_data.shape[0] == 4 * 8
    assert np.all(nb.neighbor_number[:4] == 12)


def test_system_cache_semantics():
    """Lazy attributes, reuse policy and invalidation (reference tests/test_system.py:188-215, 329-353)."""
    import mdapy_b200 as mp

    pos, box = H.fcc(3.615, 6)
    s = mp.System(pos=H.rattle(pos, 0.03, 1), box=box)
    assert not hasattr(s, "verlet_list") and not hasattr(s, "rc")
    s.build_neighbor(5.0, max_neigh=60)
    assert hasattr(s, "verlet_list") and s.rc == 5.0
    assert s.verlet_list.shape == (s.N, 60) and s.distance_list.shape == (s.N, 60)
    before = s.verlet_list.copy()
    s.cal_centro_symmetry_parameter(12)          # sorts the cached list in place (Appendix D.3)
    assert s.verlet_list.shape == (s.N, 60)
    assert np.all(np.diff(s.distance_list[:, :12], axis=1) >= 0)
    assert not np.array_equal(before, s.verlet_list)
    assert s.rc == 5.0
    s.cal_ackland_jones_analysis()               # reuses the same cached list (>= 14 neighbours)
    assert s.verlet_list.shape == (s.N, 60)
    s2 = mp.System(pos=pos, box=box)
    s2.cal_centro_symmetry_parameter(12)         # no list: kNN path, rc stays unset (system.py:1262-1263)
    assert s2.verlet_list.shape == (s2.N, 12) and not hasattr(s2, "rc")
    assert np.all(s2.neighbor_number == 12)
    s2.box = mp.Box(box * 1.01)                  # box change invalidates everything (system.py:232-245)
    assert not hasattr(s2, "verlet_list")
    # user-assigned list is honoured
    s3 = mp.System(pos=pos, box=box)
    nb = mp.Neighbor(3.0, mp.Box(box), s3.data)
    nb.compute()
    s3.verlet_list, s3.distance_list, s3.neighbor_number = nb.verlet_list, nb.distance_list, nb.neighbor_number
    s3.rc = 3.0
    s3.cal_centro_symmetry_parameter(12)
    assert np.allclose(np.asarray(s3.data["csp"]), 0.0, atol=1e-10)


def test_staged_upload_of_pageable_columns_round_trips():
    """csrc/staging.cu: pageable NumPy columns go up through page-locked ring buffers in 4 MiB slices from several
    host threads; page-locked columns go up directly.  Both must leave the same bytes on the device (odd sizes, a
    last partial slice, non-contiguous views made contiguous by the binding)."""
    import os

    from mdapy_b200 import _lib as L
    from mdapy_b200.builders import fetch_positions
    from mdapy_b200.device import DeviceSystem

    rng = np.random.default_rng(5)
    N = 3_000_001                                     # 24 MB per column: 5 full slices + a partial one
    pos = rng.random((N, 3)) * 50.0
    box, o, b = np.diag([50.0] * 3), np.zeros(3), np.array([1, 1, 1], np.int32)
    cols = [np.ascontiguousarray(pos[:, k]) for k in range(3)]
    for threads in ("1", "3", "8"):
        os.environ["MDB_UPLOAD_THREADS"] = threads
        ds = DeviceSystem(0)
        ds.set_atoms(*cols, box, o, b)
        back = fetch_positions(ds)
        assert all(np.array_equal(np.asarray(g), c) for g, c in zip(back, cols)), threads
        ds.close()
    os.environ.pop("MDB_UPLOAD_THREADS")
    pinned = [L.result_empty(N, np.float64) for _ in range(3)]
    for dst, src in zip(pinned, cols):
        dst[:] = src
    ds = DeviceSystem(0)
    ds.set_atoms(*pinned, box, o, b)
    assert all(np.array_equal(np.asarray(g), c) for g, c in zip(fetch_positions(ds), cols))
    ds.close()
