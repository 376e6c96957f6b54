"""CPU tests: pin the oracle.

1. oracle/_ref (reference C++ compiled unmodified) and oracle/port (our C restatement) each
   reproduce the reference's golden fixtures through the orchestration restatement
   (oracle/pipeline.py) -- the same assertions and tolerances the reference's own tests use
   (tests/test_common_neighbor_analysis.py:19-29, test_centro_symmetry_parameter.py:18-27,
   test_ackland_jones_analysis.py:15-25, test_polyhedral_template_matching.py:21-31,
   test_steinhardt_bond_orientation.py:25-47, test_radial_distribution_function.py:11-24).
2. port == _ref bit for bit on seeded inputs (skipped where _ref is not prebuilt).
"""
import glob
import os
from pathlib import Path

import numpy as np
import pytest

import helpers as H
from oracle import pipeline as P
from oracle import port, ref

GOLD = Path(__file__).resolve().parent / "golden"
SA = sorted(glob.glob(str(GOLD / "sa_*.npz")))
BACKENDS = [pytest.param(port, id="port", marks=pytest.mark.skipif(not port.available(), reason="port not built")),
            pytest.param(ref, id="ref", marks=pytest.mark.skipif(not ref.available(), reason="_ref not prebuilt"))]


def _load(path):
    d = np.load(path)
    return d, P.Frame(d["pos"], d["box"], d["boundary"])


@pytest.mark.parametrize("K", BACKENDS)
@pytest.mark.parametrize("path", SA, ids=[Path(p).stem[3:] for p in SA])
def test_golden_labels_and_csp(K, path):
    d, fr = _load(path)
    if "cna" in d.files:
        assert np.array_equal(P.cal_cna(K, fr, float(d["cna_cutoff"])), d["cna"])
    assert np.array_equal(P.cal_aja(K, fr), d["aja"])
    assert np.array_equal(P.cal_ids(K, fr), d["ids"])
    assert np.array_equal(P.cal_cna(K, fr, None), d["ref_acna"])
    if "csp" in d.files:
        got = P.cal_csp(K, fr, int(d["csp_num_neighbors"]))
        assert np.allclose(got, d["csp"], atol=1e-6, rtol=1e-6)


@pytest.mark.parametrize("K", BACKENDS)
@pytest.mark.parametrize("path", [p for p in SA if "q6" in np.load(p).files],
                         ids=[Path(p).stem[3:] for p in SA if "q6" in np.load(p).files])
def test_golden_steinhardt(K, path):
    d, fr = _load(path)
    rc = float(d["ql_cutoff"])
    r = P.cal_steinhardt(K, fr, [4, 6], rc=rc)
    assert np.allclose(r["qnarray"][:, 0], d["q4"], atol=1e-6, rtol=1e-6)
    assert np.allclose(r["qnarray"][:, 1], d["q6"], atol=1e-6, rtol=1e-6)
    ra = P.cal_steinhardt(K, fr, [4, 6], rc=rc, average=True)
    assert np.allclose(ra["qnarray"][:, 0], d["q4_avg"], atol=1e-6, rtol=1e-6)
    assert np.allclose(ra["qnarray"][:, 1], d["q6_avg"], atol=1e-6, rtol=1e-6)
    # w_l / w_l-hat pinned on vectors produced by the compiled reference
    rw = P.cal_steinhardt(K, fr, [4, 6, 8], rc=rc, wl=True, wlhat=True)
    assert np.allclose(rw["qnarray"], d["ref_q468_wl_wlhat"], atol=1e-12, rtol=1e-9)


@pytest.mark.skipif(not ref.available(), reason="_ref not prebuilt")
@pytest.mark.parametrize("path", SA, ids=[Path(p).stem[3:] for p in SA])
def test_golden_ptm_reference(path):
    d, fr = _load(path)
    out, ind = P.cal_ptm(ref, fr, "fcc-hcp-bcc", 0.1)
    assert np.array_equal(out[:, 0].astype(np.int32), d["ptm"])
    assert np.array_equal(out.view(np.int64), d["ref_ptm_output"].view(np.int64))
    assert np.array_equal(ind, d["ref_ptm_indices"])


@pytest.mark.parametrize("K", BACKENDS)
def test_golden_rdf(K):
    d = np.load(GOLD / "rdf_alcrni.npz")
    fr = P.Frame(d["pos"], d["box"], d["boundary"])
    res = P.cal_rdf(K, fr, float(d["cutoff"]), int(d["nbins"]), type_list=d["element"])
    el = [str(e) for e in d["elements"]]
    assert res["elements"] == sorted(el)
    for i in range(len(el)):
        for j in range(i, len(el)):
            a, b = sorted((res["elements"].index(el[i]), res["elements"].index(el[j])))
            assert np.allclose(res["g_partial"][(a, b)], d["g"][i, j], atol=1e-6)
    # streaming == list path (tests/test_rdf_streaming.py)
    res2 = P.cal_rdf(K, fr, float(d["cutoff"]), int(d["nbins"]), type_list=d["element"], streaming=True)
    assert np.allclose(res2["g_total"], res["g_total"], atol=1e-9)


@pytest.mark.parametrize("K", BACKENDS)
def test_known_answers(K):
    """Perfect-crystal invariants the reference tests assert (test_common_neighbor_analysis.py:32-45,
    test_centro_symmetry_parameter.py:30-34, test_steinhardt_bond_orientation.py:50-59)."""
    p, b = H.fcc(3.615, 5)
    fr = P.Frame(p, b)
    assert np.all(P.cal_cna(K, fr, 3.615 * 0.8536) == 1)
    assert np.all(P.cal_cna(K, fr, None) == 1)
    assert np.all(P.cal_aja(K, fr) == 1)
    assert np.allclose(P.cal_csp(K, fr, 12), 0.0, atol=1e-10)
    q = P.cal_steinhardt(K, fr, [4, 6], nnn=12)["qnarray"]
    assert np.allclose(q[:, 0], 0.190941, atol=1e-5) and np.allclose(q[:, 1], 0.574524, atol=1e-5)
    p2, b2 = H.bcc(2.8665, 6)
    fr2 = P.Frame(p2, b2)
    assert np.all(P.cal_cna(K, fr2, None) == 3)
    assert np.all(P.cal_aja(K, fr2) == 3)


# ------------------------------------------------------------------ port == reference, bit for bit
def _seeded():
    out = []
    p, b = H.fcc(3.615, 5)
    out.append(("fcc_rattled", H.rattle(p, 0.06, 0), b, [1, 1, 1], 3.3))
    out.append(("fcc_hot_slab", H.rattle(p, 0.3, 1), b, [1, 1, 0], 4.1))
    ps, bs = H.shear(H.rattle(p, 0.05, 2), b, xy=0.25, xz=-0.1, yz=0.3)
    out.append(("triclinic", ps, bs, [1, 1, 1], 3.4))
    g, bg = H.random_gas(1500, 24.0, 3)
    out.append(("gas_open", g, bg, [0, 1, 0], 3.5))
    return out


@pytest.mark.skipif(not (ref.available() and port.available()), reason="needs both checkers")
@pytest.mark.parametrize("case", _seeded(), ids=[c[0] for c in _seeded()])
def test_port_equals_reference(case):
    _, pos, box, bnd, rc = case
    x, y, z = (np.ascontiguousarray(pos[:, k]) for k in range(3))
    o = np.zeros(3)
    rv, rd, rn = ref.build_neighbor_auto(x, y, z, box, o, bnd, rc)
    pv, pd, pn = port.build_neighbor_auto(x, y, z, box, o, bnd, rc)
    assert np.array_equal(rv, pv) and np.array_equal(rn, pn) and np.array_equal(rd.view(np.int64), pd.view(np.int64))
    assert np.array_equal(ref.fcna(x, y, z, box, o, bnd, rv, rn, rc), port.fcna(x, y, z, box, o, bnd, pv, pn, rc))
    # kNN: distances bit-identical (index ties may differ)
    for k in (12, 14):
        ri, rdd = ref.knn(x, y, z, box, o, bnd, k)
        pi, pdd = port.knn(x, y, z, box, o, bnd, k)
        assert np.array_equal(rdd.view(np.int64), pdd.view(np.int64))
        uniq = np.all(np.diff(rdd, axis=1) > 0, axis=1)
        assert np.array_equal(ri[uniq], pi[uniq])
    ri, rdd = ref.knn(x, y, z, box, o, bnd, 14)
    assert np.array_equal(ref.acna(x, y, z, box, o, bnd, ri), port.acna(x, y, z, box, o, bnd, ri))
    assert np.array_equal(ref.aja(x, y, z, box, o, bnd, ri, rdd), port.aja(x, y, z, box, o, bnd, ri, rdd))
    a, b = ref.csp(x, y, z, box, o, bnd, ri, 12), port.csp(x, y, z, box, o, bnd, ri, 12)
    assert np.array_equal(a.view(np.int64), b.view(np.int64))
    rs = ref.sort_verlet_by_distance
    if rn.min() >= 6:
        v1, d1, v2, d2 = rv.copy(), rd.copy(), pv.copy(), pd.copy()
        ref.sort_verlet_by_distance(v1, d1, 6)
        port.sort_verlet_by_distance(v2, d2, 6)
        assert np.array_equal(v1, v2) and np.array_equal(d1, d2)
    for kw in (dict(), dict(average=True), dict(wl=True, wlhat=True)):
        q1 = ref.get_sq(x, y, z, box, o, bnd, rv, rd, rn, [4, 6], rc=rc, **kw)
        q2 = port.get_sq(x, y, z, box, o, bnd, pv, pd, pn, [4, 6], rc=rc, **kw)
        for u, w in zip(q1, q2):
            assert np.array_equal(np.nan_to_num(u).view(np.int64), np.nan_to_num(w).view(np.int64))
    q1 = ref.get_sq(x, y, z, box, o, bnd, rv, rd, rn, [4, 6], rc=rc)
    s1 = ref.solid_liquid(1, np.ascontiguousarray(q1[0][:, 1]), rv, rd, rn, q1[1], q1[2], 0.7, 7, rc=rc)
    s2 = port.solid_liquid(1, np.ascontiguousarray(q1[0][:, 1]), rv, rd, rn, q1[1], q1[2], 0.7, 7, rc=rc)
    assert np.array_equal(s1[0], s2[0]) and np.array_equal(s1[1], s2[1])
    t = (np.arange(x.shape[0]) % 2).astype(np.int32)
    assert np.array_equal(ref.rdf_list(rv, rd, rn, t, 2, rc, 40), port.rdf_list(pv, pd, pn, t, 2, rc, 40))
    assert np.array_equal(ref.rdf_single(rv, rd, rn, rc, 40), port.rdf_single(pv, pd, pn, rc, 40))
    assert np.array_equal(ref.rdf_streaming(x, y, z, t, 2, box, o, bnd, rc, 40),
                          port.rdf_streaming(x, y, z, t, 2, box, o, bnd, rc, 40))
    assert np.array_equal(ref.repeat_cell(box, pos[:50], 2, 3, 2), port.repeat_cell(box, pos[:50], 2, 3, 2))


def test_port_ids_equals_reference_on_defective_diamond():
    """cna.cpp:163-287 incl. the order-dependent label sweeps, on inputs with all seven labels."""
    if not ref.available():
        pytest.skip("_ref not prebuilt")
    import helpers as H

    pc, bc = H.diamond(3.567, 4)
    ph, bh = H.hex_diamond(2.522, 5, 3, 3)
    pos = np.concatenate([pc, ph + np.array([bc[0, 0] + 0.9, 0.3, 0.2])])
    pos = H.rattle(pos[np.random.default_rng(5).permutation(len(pos))], 0.03, 6)
    box = np.diag(pos.max(0) - pos.min(0) + 4.0)
    fr = P.Frame(pos - pos.min(0) + 2.0, box, [0, 0, 0])
    a, b = P.cal_ids(ref, fr), P.cal_ids(port, fr)
    assert np.array_equal(a, b) and (np.bincount(a, minlength=7) > 0).all()


# ---------------------------------------------------------------------------------------------
# further list consumers: the reference's own fixtures
# (tests/test_common_neighbor_parameter.py:21-30, test_average_neighbor.py:14-26,
#  test_warren_cowley_parameter.py:7-23)
CNP = [p for p in SA if "cnp" in np.load(p).files]


@pytest.mark.parametrize("K", BACKENDS)
@pytest.mark.parametrize("path", CNP, ids=[Path(p).stem[3:] for p in CNP])
def test_golden_cnp(K, path):
    d, fr = _load(path)
    got = P.cal_cnp(K, fr, float(d["cnp_cutoff"]))
    assert np.allclose(got, d["cnp"], atol=1e-6, rtol=1e-6), np.abs(got - d["cnp"]).max()


@pytest.mark.parametrize("K", BACKENDS)
@pytest.mark.parametrize("name", ["rec_box_big", "tri_box_big"])
def test_golden_average_by_neighbor(K, name):
    g = np.load(GOLD / "average_neighbor.npz")
    fr = P.Frame(g[f"{name}__pos"], g[f"{name}__box"], [1, 1, 1], g[f"{name}__origin"])
    got = P.cal_average_by_neighbor(K, fr, float(g[f"{name}__cutoff"]), fr.x, True)
    assert np.allclose(got, g[f"{name}__x_ave"], atol=1e-6), np.abs(got - g[f"{name}__x_ave"]).max()


@pytest.mark.parametrize("K", BACKENDS)
def test_golden_warren_cowley(K):
    g = np.load(GOLD / "wcp_cocufenipd.npz")
    fr = P.Frame(g["pos"], g["box"], g["boundary"], g["origin"])
    t = (g["type"] - 1).astype(np.int32)
    got = P.cal_wcp(K, fr, float(g["cutoff"]), t, 5)
    assert np.allclose(got.round(2), g["wcp_rounded"]), got.round(2)


@pytest.mark.skipif(not (ref.available() and port.available()), reason="needs both checkers")
def test_port_list_consumers_equal_reference():
    pos, box = H.fcc(3.615, 6)
    pos = H.rattle(pos, 0.12, 21)
    x, y, z = (np.ascontiguousarray(pos[:, k]) for k in range(3))
    o, bnd, rc = np.zeros(3), [1, 1, 1], 3.3
    v, d, n = ref.build_neighbor_auto(x, y, z, box, o, bnd, rc)
    for r in (rc, 2.9):
        a, b = ref.cnp(x, y, z, box, o, bnd, v, d, n, r), port.cnp(x, y, z, box, o, bnd, v, d, n, r)
        assert np.array_equal(a.view(np.int64), b.view(np.int64))
    t = (np.arange(x.shape[0]) % 3).astype(np.int32)
    assert np.array_equal(ref.wcp(v, n, t, 3), port.wcp(v, n, t, 3))
    for inc in (True, False):
        a, b = ref.average_by_neighbor(2.9, v, d, n, x, inc), port.average_by_neighbor(2.9, v, d, n, x, inc)
        assert np.array_equal(a.view(np.int64), b.view(np.int64))


@pytest.mark.skipif(not (ref.available() and port.available()), reason="needs both checkers")
def test_port_cluster_equals_reference():
    g, bg = H.random_gas(1500, 30.0, 12)
    x, y, z = (np.ascontiguousarray(g[:, k]) for k in range(3))
    o, bnd = np.zeros(3), [1, 1, 1]
    v, d, n = ref.build_neighbor_auto(x, y, z, bg, o, bnd, 3.0)
    for rc in (3.0, 2.2, 1.0):
        a, ca = ref.cluster(v, n, d, rc)
        b, cb = port.cluster(v, n, d, rc)
        assert ca == cb and np.array_equal(a, b) and a.min() == 1 and a.max() == ca
    t = (np.arange(x.shape[0]) % 2 + 1).astype(np.int32)
    t1, t2, r = np.array([1, 1, 2, 2], np.int32), np.array([1, 2, 1, 2], np.int32), np.array([2.0, 2.6, 2.6, 3.0])
    fa, fb = ref.filter_by_type(v, d, n, t, t1, t2, r), port.filter_by_type(v, d, n, t, t1, t2, r)
    assert np.array_equal(fa, fb) and (fa == -1).sum() > (v == -1).sum()
    a, ca = ref.cluster(fa, n)
    b, cb = port.cluster(fb, n)
    assert ca == cb and np.array_equal(a, b)


ENT_CONFIGS = ["rec_box_big", "rec_box_small", "tri_box_big", "tri_box_small"]
ENT_MODES = {"default": dict(), "use_local_density": dict(use_local_density=True), "compute_average": dict(average_rc=4.0)}


@pytest.mark.parametrize("K", BACKENDS)
@pytest.mark.parametrize("name", ENT_CONFIGS)
@pytest.mark.parametrize("mode", list(ENT_MODES))
def test_golden_structure_entropy(K, name, mode):
    """tests/test_structure_entropy.py:15-34 (rc = 5, sigma = 0.2; the small cells are replicated)."""
    g = np.load(GOLD / "structure_entropy.npz")
    fr = P.Frame(g[f"{name}__pos"], g[f"{name}__box"], [1, 1, 1], g[f"{name}__origin"])
    got = P.cal_structure_entropy(K, fr, 5.0, 0.2, **ENT_MODES[mode])
    assert np.allclose(got, g[f"{name}__{mode}"], atol=1e-6), np.abs(got - g[f"{name}__{mode}"]).max()


@pytest.mark.skipif(not (ref.available() and port.available()), reason="needs both checkers")
def test_port_structure_entropy_equals_reference():
    pos, box = H.fcc(4.05, 5)
    pos = H.rattle(pos, 0.15, 8)
    x, y, z = (np.ascontiguousarray(pos[:, k]) for k in range(3))
    v, d, n = ref.build_neighbor_auto(x, y, z, box, np.zeros(3), [1, 1, 1], 5.0)
    vol = float(np.linalg.det(box))
    for uld in (False, True):
        a, b = ref.structure_entropy(5.0, 0.2, uld, vol, d, n), port.structure_entropy(5.0, 0.2, uld, vol, d, n)
        assert np.array_equal(a.view(np.int64), b.view(np.int64))


@pytest.mark.skipif(not (ref.available() and port.available()), reason="needs both checkers")
def test_port_atomic_temperature_equals_reference():
    pos, box = H.fcc(3.615, 6)
    pos = H.rattle(pos, 0.1, 9)
    x, y, z = (np.ascontiguousarray(pos[:, k]) for k in range(3))
    v, d, n = ref.build_neighbor_auto(x, y, z, box, np.zeros(3), [1, 1, 1], 4.0)
    rng = np.random.default_rng(4)
    vel = rng.standard_normal((3, x.shape[0])) * 3.0
    mass = rng.choice([26.98, 63.546, 58.69], x.shape[0])
    for rc in (4.0, 3.0):
        a = ref.compute_temp(v, d, vel[0], vel[1], vel[2], mass, rc)
        b = port.compute_temp(v, d, vel[0], vel[1], vel[2], mass, rc)
        assert np.array_equal(a.view(np.int64), b.view(np.int64)) and a.min() > 0


@pytest.mark.skipif(not (ref.available() and port.available()), reason="needs both checkers")
def test_port_bond_analysis_equals_reference():
    pos, box = H.fcc(3.615, 6)
    pos = H.rattle(pos, 0.12, 13)
    x, y, z = (np.ascontiguousarray(pos[:, k]) for k in range(3))
    o, bnd = np.zeros(3), [1, 1, 0]
    v, d, n = ref.build_neighbor_auto(x, y, z, box, o, bnd, 3.4)
    a, b = ref.compute_bond(x, y, z, box, o, bnd, v, d, n, 3.4, 50), port.compute_bond(x, y, z, box, o, bnd, v, d, n, 3.4, 50)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and a[0].sum() > 0 and a[1].sum() > 0
    t = (np.arange(x.shape[0]) % 2).astype(np.int32)
    rcl = np.array([[0.0, 3.0, 0.0, 3.4], [2.0, 3.4, 2.0, 3.4]])
    pl = np.array([[0, 1, 0], [1, 1, 1]], np.int32)
    assert np.array_equal(ref.compute_adf(x, y, z, box, o, bnd, v, d, n, rcl, pl, t, 36),
                          port.compute_adf(x, y, z, box, o, bnd, v, d, n, rcl, pl, t, 36))


@pytest.mark.skipif(not (ref.available() and port.available()), reason="needs both checkers")
def test_port_wrap_positions_equals_reference():
    rng = np.random.default_rng(2)
    p, b = H.fcc(3.615, 4)
    far = H.rattle(p, 0.3, 3) + rng.integers(-3, 4, p.shape) * np.diag(b)
    ps, bs = H.shear(far, b, xy=0.25, xz=-0.1, yz=0.2)
    for pos, box, origin, bnd in ((far, b, np.array([-3.0, 1.5, 0.25]), [1, 1, 1]), (ps, bs, np.zeros(3), [1, 0, 1])):
        a = [np.ascontiguousarray(pos[:, k]) for k in range(3)]
        c = [v.copy() for v in a]
        ref.wrap_positions(*a, box, origin, bnd)
        port.wrap_positions(*c, box, origin, bnd)
        for u, w in zip(a, c):
            assert np.array_equal(u.view(np.int64), w.view(np.int64))


GOLD_DIR = Path(__file__).resolve().parent / "golden"


# ---- further reference fixtures pinned on the compiled reference (round 2)
def test_ref_chill_plus_reproduces_the_upstream_fixture():
    if not ref.available():
        pytest.skip("oracle/_ref not built")
    d = np.load(GOLD_DIR / "chill_water.npz")
    pos, box, bnd, rc = d["pos"], d["box"], d["boundary"], float(d["chill_plus_cutoff"])
    x, y, z = (np.ascontiguousarray(pos[:, k]) for k in range(3))
    v, dd, n = ref.build_neighbor_auto(x, y, z, box, np.zeros(3), bnd, rc)
    assert np.array_equal(ref.chill_plus(x, y, z, box, np.zeros(3), bnd, v, dd, n, rc), d["chill_plus"])


def test_ref_planar_faults_reproduce_the_upstream_fixture():
    if not ref.available():
        pytest.skip("oracle/_ref not built")
    from oracle import pipeline as P

    d = np.load(GOLD_DIR / "fcc_planar_faults.npz")
    fr = P.Frame(d["pos"], d["box"], d["boundary"], d["origin"])
    out, ind = P.cal_ptm(ref, fr, "all", 0.1)
    got = ref.identify_sftb_fcc(out[:, 0].astype(np.int32), ind[:, 1:13], identify_esf=False)
    assert np.array_equal(got, d["pft"])
