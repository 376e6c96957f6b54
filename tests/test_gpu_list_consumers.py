"""GPU parity of the further neighbour-list consumers: common neighbour parameter,
Warren-Cowley parameter, average_by_neighbor -- device-resident path, host-pointer C ABI and the
System methods, against the oracle (bit-exact) and the reference's own fixtures."""
import glob
from pathlib import Path

import numpy as np
import pytest

import helpers as H
from oracle import checker as K
from oracle import pipeline as P

pytestmark = pytest.mark.gpu
GOLD = Path(__file__).resolve().parent / "golden"
SA = sorted(glob.glob(str(GOLD / "sa_*.npz")))
CNP = [p for p in SA if "cnp" in np.load(p).files]


def _cases():
    p, b = H.fcc(3.615, 8)
    out = [("fcc_rattled", H.rattle(p, 0.1, 3), b, [1, 1, 1], 3.2),
           ("fcc_hot_slab", H.rattle(p, 0.3, 4), b, [1, 1, 0], 3.6)]
    ps, bs = H.shear(H.rattle(p, 0.08, 5), b, xy=0.2, xz=0.1, yz=-0.15)
    out.append(("triclinic", ps, bs, [1, 1, 1], 3.3))
    g, bg = H.random_gas(2500, 30.0, 6)
    out.append(("gas", g, bg, [1, 1, 1], 4.5))
    return out


CASES = _cases()


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_device_path_bit_exact(case):
    from mdapy_b200.device import DeviceSystem

    _, pos, box, bnd, rc = case
    x, y, z = (np.ascontiguousarray(pos[:, k]) for k in range(3))
    o = np.zeros(3)
    v, d, n = K.build_neighbor_auto(x, y, z, box, o, bnd, rc)
    ds = DeviceSystem(0)
    ds.set_atoms(x, y, z, box, o, bnd)
    ds.build_neighbor(rc)
    for r in (rc, 0.9 * rc):
        ref = K.cnp(x, y, z, box, o, bnd, v, d, n, r)
        got = ds.cnp(r)
        assert np.array_equal(got.view(np.int64), ref.view(np.int64)), np.abs(got - ref).max()
    t = (np.random.default_rng(0).integers(0, 3, x.shape[0])).astype(np.int32)
    assert np.array_equal(ds.wcp(t, 3), K.wcp(v, n, t, 3))
    val = np.random.default_rng(1).normal(size=x.shape[0])
    for inc in (True, False):
        ref = K.average_by_neighbor(0.9 * rc, v, d, n, val, inc)
        got = ds.average_by_neighbor(0.9 * rc, val, inc)
        assert np.array_equal(got.view(np.int64), ref.view(np.int64))


def test_host_pointer_dropins():
    """Section A: same arguments as the reference's nanobind functions, host arrays in and out."""
    from mdapy_b200 import _lib as L

    _, pos, box, bnd, rc = CASES[0]
    x, y, z = (np.ascontiguousarray(pos[:, k]) for k in range(3))
    o = np.zeros(3)
    v, d, n = K.build_neighbor_auto(x, y, z, box, o, bnd, rc)
    b, oo, p = L.box_args(box, o, bnd)
    N, M = v.shape
    out = np.zeros(N)
    L.check(L.lib().mdb_compute_cnp(L.dptr(x), L.dptr(y), L.dptr(z), N, L.dptr(b), L.dptr(oo), L.iptr(p), L.iptr(v), M,
                                    L.dptr(d), L.iptr(n), L.dptr(out), rc, 1))
    assert np.array_equal(out.view(np.int64), K.cnp(x, y, z, box, o, bnd, v, d, n, rc).view(np.int64))
    t = (np.arange(N) % 4).astype(np.int32)
    w = np.zeros((4, 4))
    L.check(L.lib().mdb_get_wcp(L.iptr(v), N, M, L.iptr(n), L.iptr(t), 4, L.dptr(w), 1))
    assert np.array_equal(w, K.wcp(v, n, t, 4))
    ave = np.zeros(N)
    L.check(L.lib().mdb_average_by_neighbor(rc, L.iptr(v), N, M, L.dptr(d), L.iptr(n), L.dptr(x), L.dptr(ave), 1, 1))
    assert np.array_equal(ave.view(np.int64), K.average_by_neighbor(rc, v, d, n, x, True).view(np.int64))


@pytest.mark.parametrize("path", CNP, ids=[Path(p).stem[3:] for p in CNP])
def test_system_cnp_against_reference_fixture(path):
    """tests/test_common_neighbor_parameter.py:21-30 (incl. the small boxes that are replicated)."""
    import mdapy_b200 as mp

    d = np.load(path)
    system = mp.System(pos=d["pos"], box=mp.Box(d["box"], d["boundary"]))
    system.cal_common_neighbor_parameter(float(d["cnp_cutoff"]))
    got = np.asarray(system.data["cnp"])
    assert np.allclose(got, d["cnp"], atol=1e-6, rtol=1e-6), np.abs(got - d["cnp"]).max()
    fr = P.Frame(d["pos"], d["box"], d["boundary"])
    ref = P.cal_cnp(K, fr, float(d["cnp_cutoff"]))
    assert np.array_equal(got.view(np.int64), ref.view(np.int64))


def test_system_cnp_known_answers():
    """tests/test_common_neighbor_parameter.py:33-53."""
    import mdapy_b200 as mp

    a = 3.615
    s = mp.build_crystal("Cu", "fcc", a, nx=4, ny=4, nz=4)
    s.cal_common_neighbor_parameter(0.86 * a)
    assert np.allclose(np.asarray(s.data["cnp"]).max(), 0.0)
    s = mp.build_crystal("Cu", "bcc", a, nx=4, ny=4, nz=4)
    s.cal_common_neighbor_parameter(1.21 * a)
    assert np.allclose(np.asarray(s.data["cnp"]).max(), 0.0)


@pytest.mark.parametrize("name", ["rec_box_big", "tri_box_big"])
def test_system_average_by_neighbor_fixture(name):
    """tests/test_average_neighbor.py:14-26."""
    import mdapy_b200 as mp

    g = np.load(GOLD / "average_neighbor.npz")
    system = mp.System(pos=g[f"{name}__pos"], box=mp.Box(g[f"{name}__box"], [1, 1, 1], g[f"{name}__origin"]))
    system.average_by_neighbor(float(g[f"{name}__cutoff"]), "x", include_self=True)
    got = np.asarray(system.data["x_ave"])
    assert np.allclose(got, g[f"{name}__x_ave"], atol=1e-6), np.abs(got - g[f"{name}__x_ave"]).max()


def test_system_warren_cowley_fixture():
    """tests/test_warren_cowley_parameter.py:7-23 (8788-atom CoCuFeNiPd, non-zero box origin)."""
    import mdapy_b200 as mp

    g = np.load(GOLD / "wcp_cocufenipd.npz")
    pos = g["pos"]
    system = mp.System(data={"x": pos[:, 0].copy(), "y": pos[:, 1].copy(), "z": pos[:, 2].copy(), "type": g["type"]},
                       box=mp.Box(g["box"], g["boundary"], g["origin"]))
    wcp = system.cal_warren_cowley_parameter(rc=float(g["cutoff"]))
    assert np.allclose(wcp.WCP.round(2), g["wcp_rounded"]), wcp.WCP.round(2)
    fr = P.Frame(pos, g["box"], g["boundary"], g["origin"])
    assert np.array_equal(wcp.WCP, P.cal_wcp(K, fr, float(g["cutoff"]), (g["type"] - 1).astype(np.int32), 5))


# ---------------------------------------------------------------------------------------------
# cluster analysis (src/cluster.cpp): ids numbered like the reference's serial flood
def _gas(n, L, seed):
    g, bg = H.random_gas(n, L, seed)
    return np.ascontiguousarray(g[:, 0]), np.ascontiguousarray(g[:, 1]), np.ascontiguousarray(g[:, 2]), bg


@pytest.mark.parametrize("rc_list,rc", [(3.0, 3.0), (3.0, 2.1), (3.0, 0.8), (4.5, 4.5)])
def test_cluster_ids_equal_reference(rc_list, rc):
    from mdapy_b200.device import DeviceSystem

    x, y, z, box = _gas(4000, 40.0, 14)
    o, bnd = np.zeros(3), [1, 1, 0]
    v, d, n = K.build_neighbor_auto(x, y, z, box, o, bnd, rc_list)
    ref, cnt = K.cluster(v, n, d, rc)
    ds = DeviceSystem(0)
    ds.set_atoms(x, y, z, box, o, bnd)
    ds.build_neighbor(rc_list)
    got, c = ds.cluster(rc)
    assert c == cnt and np.array_equal(got, ref)


def test_cluster_type_pair_cutoffs_and_host_api():
    import ctypes as C

    from mdapy_b200 import _lib as L
    from mdapy_b200.cluster_analysis import ClusterAnalysis, type_pair_table
    from mdapy_b200.device import DeviceSystem

    x, y, z, box = _gas(3000, 36.0, 15)
    o, bnd = np.zeros(3), [1, 1, 1]
    types = (np.random.default_rng(2).integers(1, 3, x.shape[0])).astype(np.int32)
    rcd = {"1-1": 2.0, "1-2": 2.7, "2-2": 3.1}
    v, d, n = K.build_neighbor_auto(x, y, z, box, o, bnd, 3.1)
    t1, t2, r = type_pair_table(rcd)
    fv = K.filter_by_type(v, d, n, types, t1, t2, r)
    ref, cnt = K.cluster(fv, n)
    # device-resident path
    ds = DeviceSystem(0)
    ds.set_atoms(x, y, z, box, o, bnd)
    ds.build_neighbor(3.1)
    got, c = ds.cluster(0.0, types, t1, t2, r)
    assert c == cnt and np.array_equal(got, ref)
    # host-pointer path through the mirrored class (filter_by_type + get_cluster_by_bond)
    ca = ClusterAnalysis(rcd, v, d, n, types)
    ca.compute()
    assert np.array_equal(ca.verlet_list, fv) and ca.cluster_number == cnt and np.array_equal(ca.particleClusters, ref)
    ca = ClusterAnalysis(2.4, v, d, n)
    ca.compute()
    ref2, cnt2 = K.cluster(v, n, d, 2.4)
    assert ca.cluster_number == cnt2 and np.array_equal(ca.particleClusters, ref2)


def test_system_cluster_analysis():
    import mdapy_b200 as mp

    x, y, z, box = _gas(3000, 36.0, 16)
    types = (np.arange(x.shape[0]) % 2 + 1).astype(np.int32)
    system = mp.System(data={"x": x, "y": y, "z": z, "type": types}, box=mp.Box(box))
    system.cal_cluster_analysis(2.5)
    v, d, n = K.build_neighbor_auto(x, y, z, box, np.zeros(3), [1, 1, 1], 2.5)
    ref, cnt = K.cluster(v, n, d, 2.5)
    assert np.array_equal(np.asarray(system.data["cluster_id"]), ref)
    system.cal_cluster_analysis({"1-1": 1.8, "1-2": 2.5, "2-2": 2.2})      # reuses the cached rc = 2.5 list
    from mdapy_b200.cluster_analysis import type_pair_table

    t1, t2, r = type_pair_table({"1-1": 1.8, "1-2": 2.5, "2-2": 2.2})
    ref, cnt = K.cluster(K.filter_by_type(v, d, n, types, t1, t2, r), n)
    assert np.array_equal(np.asarray(system.data["cluster_id"]), ref)


# ---------------------------------------------------------------------------------------------
# structure entropy (src/structure_entropy.cpp): exp / log come from the device's libm, so the bar is the
# north star's floating-point tolerance (1e-6 relative) -- observed agreement is ~1e-14
@pytest.mark.parametrize("uld", [False, True])
def test_structure_entropy_device_and_host_api(uld):
    from mdapy_b200 import _lib as L
    from mdapy_b200.device import DeviceSystem

    p, b = H.fcc(4.05, 8)
    pos = H.rattle(p, 0.2, 17)
    x, y, z = (np.ascontiguousarray(pos[:, k]) for k in range(3))
    o, bnd, rc, sigma = np.zeros(3), [1, 1, 1], 5.0, 0.2
    v, d, n = K.build_neighbor_auto(x, y, z, b, o, bnd, rc)
    vol = float(np.linalg.det(b))
    ref = K.structure_entropy(rc, sigma, uld, vol, d, n)
    ds = DeviceSystem(0)
    ds.set_atoms(x, y, z, b, o, bnd)
    ds.build_neighbor(rc)
    ent, ave = ds.structure_entropy(rc, sigma, uld, vol, 4.0)
    assert np.allclose(ent, ref, rtol=1e-6, atol=0) and np.abs(ent / ref - 1).max() < 1e-11
    ref_ave = K.average_by_neighbor(4.0, v, d, n, ref, True)
    assert np.allclose(ave, ref_ave, rtol=1e-6, atol=0)
    out = np.zeros(x.shape[0])
    L.check(L.lib().mdb_calculate_structure_entropy(rc, sigma, int(uld), vol, L.dptr(d), d.shape[0], d.shape[1], L.iptr(n),
                                                    L.dptr(out), 1))
    assert np.array_equal(out.view(np.int64), np.asarray(ent).view(np.int64))


@pytest.mark.parametrize("name", ["rec_box_big", "rec_box_small", "tri_box_big", "tri_box_small"])
@pytest.mark.parametrize("mode", ["default", "use_local_density", "compute_average"])
def test_system_structure_entropy_fixture(name, mode):
    """tests/test_structure_entropy.py:15-34."""
    import mdapy_b200 as mp

    g = np.load(GOLD / "structure_entropy.npz")
    system = mp.System(pos=g[f"{name}__pos"], box=mp.Box(g[f"{name}__box"], [1, 1, 1], g[f"{name}__origin"]))
    if mode == "compute_average":
        system.cal_structure_entropy(5.0, 0.2, False, average_rc=4.0)
        got = np.asarray(system.data["entropy_ave"])
    elif mode == "use_local_density":
        system.cal_structure_entropy(5.0, 0.2, True)
        got = np.asarray(system.data["entropy"])
    else:
        system.cal_structure_entropy(5.0, 0.2, False)
        got = np.asarray(system.data["entropy"])
    exp = g[f"{name}__{mode}"]
    assert np.allclose(got, exp, atol=1e-6), np.abs(got - exp).max()


# ---------------------------------------------------------------------------------------------
# atomic temperature (src/atomic_temperature.cpp)
def test_atomic_temperature_bit_exact():
    import mdapy_b200 as mp
    from mdapy_b200 import _lib as L
    from mdapy_b200.device import DeviceSystem

    p, b = H.fcc(3.615, 8)
    pos = H.rattle(p, 0.1, 19)
    x, y, z = (np.ascontiguousarray(pos[:, k]) for k in range(3))
    o, bnd, rc = np.zeros(3), [1, 1, 1], 4.0
    v, d, n = K.build_neighbor_auto(x, y, z, b, o, bnd, rc)
    rng = np.random.default_rng(5)
    vel = rng.standard_normal((3, x.shape[0])) * 2.5           # A/ps
    mass = rng.choice([26.98, 63.546], x.shape[0])
    ds = DeviceSystem(0)
    ds.set_atoms(x, y, z, b, o, bnd)
    ds.build_neighbor(rc)
    for r in (rc, 3.0):
        ref = K.compute_temp(v, d, vel[0], vel[1], vel[2], mass, r)
        got = ds.atomic_temperature(vel[0], vel[1], vel[2], mass, r)
        assert np.array_equal(got.view(np.int64), ref.view(np.int64))
    out = np.zeros(x.shape[0])
    L.check(L.lib().mdb_compute_temp(L.iptr(v), v.shape[0], v.shape[1], L.dptr(d), L.dptr(np.ascontiguousarray(vel[0])),
                                     L.dptr(np.ascontiguousarray(vel[1])), L.dptr(np.ascontiguousarray(vel[2])),
                                     L.dptr(mass), L.dptr(out), rc, 1))
    assert np.array_equal(out.view(np.int64), K.compute_temp(v, d, vel[0], vel[1], vel[2], mass, rc).view(np.int64))
    # System method: velocities in A/fs, scaled by 1e3 * factor like the reference's wrapper
    system = mp.System(data={"x": x, "y": y, "z": z, "vx": vel[0] * 1e-3, "vy": vel[1] * 1e-3, "vz": vel[2] * 1e-3,
                             "amass": mass}, box=mp.Box(b))
    system.cal_atomic_temperature(rc)
    ref = K.compute_temp(v, d, vel[0] * 1e-3 * 1e3 * 1.0, vel[1] * 1e-3 * 1e3 * 1.0, vel[2] * 1e-3 * 1e3 * 1.0, mass, rc)
    assert np.array_equal(np.asarray(system.data["atomic_temp"]).view(np.int64), ref.view(np.int64))


# ---------------------------------------------------------------------------------------------
# bond-length / bond-angle histograms and the angular distribution function (src/bond_analysis.cpp).
# Rattled inputs: no angle sits within rounding of a bin edge, so the device acos cannot move a count.
@pytest.mark.parametrize("case", CASES[:3], ids=[c[0] for c in CASES[:3]])
def test_bond_analysis_and_adf_exact(case):
    from mdapy_b200 import _lib as L
    from mdapy_b200.device import DeviceSystem

    _, pos, box, bnd, rc = case
    x, y, z = (np.ascontiguousarray(pos[:, k]) for k in range(3))
    o = np.zeros(3)
    v, d, n = K.build_neighbor_auto(x, y, z, box, o, bnd, rc)
    ds = DeviceSystem(0)
    ds.set_atoms(x, y, z, box, o, bnd)
    ds.build_neighbor(rc)
    for nbin in (40, 181):
        rl, ra = K.compute_bond(x, y, z, box, o, bnd, v, d, n, rc, nbin)
        gl, ga = ds.bond_analysis(rc, nbin)
        assert np.array_equal(gl, rl) and np.array_equal(ga, ra) and ra.sum() > 0
    t = (np.random.default_rng(3).integers(0, 2, x.shape[0])).astype(np.int32)
    rcl = np.array([[0.0, 0.9 * rc, 0.0, rc], [0.5 * rc, rc, 0.5 * rc, rc], [0.0, rc, 0.0, 0.8 * rc]])
    pl = np.array([[0, 1, 0], [1, 1, 1], [1, 0, 1]], np.int32)
    ref = K.compute_adf(x, y, z, box, o, bnd, v, d, n, rcl, pl, t, 60)
    assert np.array_equal(ds.adf(rcl, pl, t, 60), ref) and ref.sum() > 0
    # host-pointer drop-ins accumulate into the caller's histograms
    b, oo, p = L.box_args(box, o, bnd)
    N, M = v.shape
    bl, ba = np.ones(40, np.int32), np.zeros(40, np.int32)
    L.check(L.lib().mdb_compute_bond(L.dptr(x), L.dptr(y), L.dptr(z), N, L.dptr(b), L.dptr(oo), L.iptr(p), L.iptr(v), M,
                                     L.dptr(d), L.iptr(n), L.iptr(bl), L.iptr(ba), rc / 40, 180.0 / 40, rc, 40, 1))
    rl, ra = K.compute_bond(x, y, z, box, o, bnd, v, d, n, rc, 40)
    assert np.array_equal(bl, rl + 1) and np.array_equal(ba, ra)


def test_system_bond_analysis_and_adf():
    import mdapy_b200 as mp

    p, b = H.fcc(3.615, 8)
    pos = H.rattle(p, 0.1, 23)
    x, y, z = (np.ascontiguousarray(pos[:, k]) for k in range(3))
    el = np.array(["Cu", "Ni"], dtype=object)[np.random.default_rng(6).integers(0, 2, x.shape[0])]
    system = mp.System(data={"x": x, "y": y, "z": z, "element": el}, box=mp.Box(b))
    ba = system.cal_bond_analysis(3.2, 64)
    v, d, n = K.build_neighbor_auto(x, y, z, b, np.zeros(3), [1, 1, 1], 3.2)
    rl, ra = K.compute_bond(x, y, z, b, np.zeros(3), [1, 1, 1], v, d, n, 3.2, 64)
    assert np.array_equal(ba.bond_length_distribution, rl) and np.array_equal(ba.bond_angle_distribution, ra)
    assert ba.r_angle.shape == (64,) and abs(ba.r_length[0] - 3.2 / 128) < 1e-12
    adf = system.cal_angular_distribution_function({"Cu-Ni-Ni": [0.0, 3.0, 0.0, 3.2], "Ni-Cu-Ni": [0.0, 3.2, 0.0, 3.2]}, 45)
    t = np.array([{"Cu": 0, "Ni": 1}[e] for e in el], np.int32)
    ref = K.compute_adf(x, y, z, b, np.zeros(3), [1, 1, 1], v, d, n, np.array([[0.0, 3.0, 0.0, 3.2], [0.0, 3.2, 0.0, 3.2]]),
                        np.array([[0, 1, 1], [1, 0, 1]], np.int32), t, 45)
    assert np.array_equal(adf.bond_angle_distribution, ref)
