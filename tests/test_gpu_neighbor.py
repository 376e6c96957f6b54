"""GPU parity: cut-off neighbour build + sort + CNA/CSP/AJA kernels against the oracle.

Bit-exact bar: counts, index rows (including ORDER), f64 distances, labels.
Mirrors the cases of the reference's tests/test_neighbor_cutoff.py (ortho /
triclinic / mixed PBC / zero neighbours / exact-at-cutoff / max_neigh).
"""
import numpy as np
import pytest

import helpers as H
from oracle import checker as K

pytestmark = pytest.mark.gpu


def _dev():
    from mdapy_b200.device import DeviceSystem

    return DeviceSystem(0)


def _cases():
    out = []
    p, b = H.fcc(3.615, 6)
    out.append(("fcc6", p, b, [1, 1, 1], 3.615 * 0.8536))
    out.append(("fcc6_rc5", p, b, [1, 1, 1], 5.0))
    out.append(("fcc6_rattled", H.rattle(p, 0.05, 0), b, [1, 1, 1], 3.615 * 0.8536))
    out.append(("fcc6_hot_unwrapped", H.rattle(p, 0.4, 1), b, [1, 1, 1], 4.2))
    out.append(("fcc6_slab", H.rattle(p, 0.05, 2), b, [1, 1, 0], 3.4))
    out.append(("fcc6_wire", H.rattle(p, 0.05, 3), b, [1, 0, 0], 3.4))
    out.append(("fcc6_open", H.rattle(p, 0.05, 4), b, [0, 0, 0], 3.4))
    ps, bs = H.shear(H.rattle(p, 0.05, 5), b, xy=0.2, xz=0.1, yz=-0.15)
    out.append(("fcc6_triclinic", ps, bs, [1, 1, 1], 3.4))
    ps, bs = H.shear(H.rattle(p, 0.05, 6), b, xy=0.6, xz=0.0, yz=0.45)
    out.append(("fcc6_tilted_mixed", ps, bs, [1, 0, 1], 3.0))
    p2, b2 = H.bcc(2.8665, 7)
    out.append(("bcc7", p2, b2, [1, 1, 1], 2.8665 * 1.2))
    g, bg = H.random_gas(3000, 30.0, 7)
    out.append(("gas", g, bg, [1, 1, 1], 4.0))
    g2, bg2 = H.random_gas(200, 60.0, 8)
    out.append(("sparse_zero_neigh", g2, bg2, [1, 1, 1], 0.5))
    # triclinic frames large enough for the cell-tile kernels (>= 7 cells per periodic axis): sheared, strongly
    # tilted with an open axis, and an origin off zero
    p12, b12 = H.fcc(3.615, 12)
    ps, bs = H.shear(H.rattle(p12, 0.06, 21), b12, xy=0.2, xz=0.1, yz=-0.15)
    out.append(("fcc12_triclinic", ps, bs, [1, 1, 1], 3.615 * 0.8536))
    ps, bs = H.shear(H.rattle(p12, 0.25, 22), b12, xy=0.55, xz=-0.3, yz=0.4)
    out.append(("fcc12_tilted_hot_mixed", ps, bs, [1, 0, 1], 3.3))
    ps, bs = H.shear(H.rattle(p12, 0.06, 23), b12, xy=-0.25, xz=0.0, yz=0.2)
    out.append(("fcc12_triclinic_rc5", ps, bs, [1, 1, 1], 5.0))
    # far outside the box on every axis (unwrapped trajectory)
    out.append(("fcc6_far_images", H.rattle(p, 0.05, 9) + np.array([3, -2, 5]) * np.diag(b), b, [1, 1, 1], 3.2))
    return out


CASES = _cases()


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
@pytest.mark.parametrize("max_neigh", [None, 60])
def test_neighbor_rows_bit_exact(case, max_neigh):
    name, pos, box, boundary, rc = case
    x, y, z = (np.ascontiguousarray(pos[:, k]) for k in range(3))
    origin = np.zeros(3)
    if max_neigh is None:
        rv, rd, rn = K.build_neighbor_auto(x, y, z, box, origin, boundary, rc)
    else:
        rv, rd, rn = K.build_neighbor(x, y, z, box, origin, boundary, rc, max_neigh)
    ds = _dev()
    ds.set_atoms(x, y, z, box, origin, boundary)
    M, mx = ds.build_neighbor(rc, max_neigh)
    v, d, n = ds.fetch_neighbor()
    assert M == rv.shape[1]
    assert np.array_equal(n, rn)
    assert np.array_equal(v, rv), "row order / membership differs from the reference"
    assert np.array_equal(d.view(np.int64), rd.view(np.int64)), "distances are not bit-identical"


def test_max_neigh_too_small_counts_true_number():
    p, b = H.fcc(3.615, 5)
    x, y, z = (np.ascontiguousarray(p[:, k]) for k in range(3))
    rv, rd, rn = K.build_neighbor(x, y, z, b, np.zeros(3), [1, 1, 1], 3.615 * 0.8536, 5)
    ds = _dev()
    ds.set_atoms(x, y, z, b, np.zeros(3), [1, 1, 1])
    M, mx = ds.build_neighbor(3.615 * 0.8536, 5)
    v, d, n = ds.fetch_neighbor()
    assert mx == 12 and np.all(n == 12)
    assert np.array_equal(v, rv) and np.array_equal(d, rd)


@pytest.mark.parametrize("case", CASES[:10], ids=[c[0] for c in CASES[:10]])
def test_sort_and_descriptors(case):
    name, pos, box, boundary, rc = case
    x, y, z = (np.ascontiguousarray(pos[:, k]) for k in range(3))
    origin = np.zeros(3)
    rv, rd, rn = K.build_neighbor_auto(x, y, z, box, origin, boundary, rc)
    ds = _dev()
    ds.set_atoms(x, y, z, box, origin, boundary)
    ds.build_neighbor(rc)
    # fixed-cutoff CNA on the raw list
    ref_cna = K.fcna(x, y, z, box, origin, boundary, rv, rn, rc)
    assert np.array_equal(ds.fcna(rc), ref_cna)
    kmin = int(rn.min())
    for k in (12, 14):
        if kmin < k:
            continue
        K.sort_verlet_by_distance(rv, rd, k)
        ds.sort_neighbor(k)
        v, d, n = ds.fetch_neighbor()
        assert np.array_equal(v, rv) and np.array_equal(d.view(np.int64), rd.view(np.int64))
    if kmin >= 12:
        ref_csp = K.csp(x, y, z, box, origin, boundary, rv, 12)
        got = ds.csp(12)
        assert np.array_equal(got.view(np.int64), ref_csp.view(np.int64)), np.abs(got - ref_csp).max()
    if kmin >= 14:
        assert np.array_equal(ds.aja(), K.aja(x, y, z, box, origin, boundary, rv, rd))
        assert np.array_equal(ds.acna(), K.acna(x, y, z, box, origin, boundary, rv))


def test_config1_fcc_32k():
    """BASELINE config 1: 32,000-atom FCC Cu, rc = 3.615*0.8536 -> nn == 12, cna == 1."""
    p, b = H.fcc(3.615, 20)
    x, y, z = (np.ascontiguousarray(p[:, k]) for k in range(3))
    rc = 3.615 * 0.8536
    ds = _dev()
    ds.set_atoms(x, y, z, b, np.zeros(3), [1, 1, 1])
    M, mx = ds.build_neighbor(rc)
    v, d, n = ds.fetch_neighbor()
    assert M == 12 and np.all(n == 12)
    cna = ds.fcna(rc)
    assert np.all(cna == 1)
    rv, rd, rn = K.build_neighbor_auto(x, y, z, b, np.zeros(3), [1, 1, 1], rc)
    assert np.array_equal(v, rv) and np.array_equal(d.view(np.int64), rd.view(np.int64))
    for sigma, seed in ((0.05, 0), (0.20, 1)):
        pr = H.rattle(p, sigma, seed)
        x, y, z = (np.ascontiguousarray(pr[:, k]) for k in range(3))
        rv, rd, rn = K.build_neighbor_auto(x, y, z, b, np.zeros(3), [1, 1, 1], rc)
        ds.set_atoms(x, y, z, b, np.zeros(3), [1, 1, 1])
        ds.build_neighbor(rc)
        v, d, n = ds.fetch_neighbor()
        assert np.array_equal(v, rv) and np.array_equal(n, rn) and np.array_equal(d.view(np.int64), rd.view(np.int64))
        assert np.array_equal(ds.fcna(rc), K.fcna(x, y, z, b, np.zeros(3), [1, 1, 1], rv, rn, rc))


# ---------------------------------------------------------------------------------------------
# Cell-tile kernel (neighbor_tiled.cu): frames large enough to take it, exercising every branch --
# tile shapes from sparse to dense, non-zero origin (wrap is not the identity bit pattern), unwrapped
# atoms (periodic-image path), open axes, 16-byte and scalar row stores, the automatic width with margin +
# compaction, overflowing max_neigh.
def _tiled_cases():
    out = []
    p, b = H.fcc(3.615, 12)                       # 43.4 A: 14 cells at rc = 3.08
    r = H.rattle(p, 0.06, 31)
    out.append(("fcc12_T4", r, b, np.zeros(3), [1, 1, 1], 3.615 * 0.8536))
    out.append(("fcc12_origin", r + np.array([-7.3, 11.1, 0.37]), b, np.array([-7.3, 11.1, 0.37]), [1, 1, 1], 3.1))
    shift = np.random.default_rng(5).integers(-2, 3, r.shape) * np.diag(b)
    out.append(("fcc12_unwrapped", r + shift, b, np.zeros(3), [1, 1, 1], 3.1))
    out.append(("fcc12_slab", r, b, np.zeros(3), [1, 1, 0], 3.3))
    out.append(("fcc12_open", r, b, np.zeros(3), [0, 0, 0], 3.3))
    out.append(("fcc12_dense_rc5", r, b, np.zeros(3), [1, 1, 1], 5.0))
    out.append(("fcc12_dense_rc6", r, b, np.zeros(3), [1, 0, 1], 6.2))
    g, bg = H.random_gas(6000, 60.0, 33)           # 0.03 atoms / A^3: sparse tiles
    out.append(("gas_sparse", g, bg, np.zeros(3), [1, 1, 1], 2.4))
    g2, bg2 = H.random_gas(20000, 50.0, 34)        # clustered cells, wide count distribution
    out.append(("gas_dense", g2, bg2, np.zeros(3), [1, 1, 1], 3.3))
    return out


TILED = _tiled_cases()


@pytest.mark.parametrize("case", TILED, ids=[c[0] for c in TILED])
@pytest.mark.parametrize("max_neigh", [None, 13, 64])
def test_tiled_kernel_rows_bit_exact(case, max_neigh):
    name, pos, box, origin, boundary, rc = case
    x, y, z = (np.ascontiguousarray(pos[:, k]) for k in range(3))
    if max_neigh is None:
        rv, rd, rn = K.build_neighbor_auto(x, y, z, box, origin, boundary, rc)
    else:
        rv, rd, rn = K.build_neighbor(x, y, z, box, origin, boundary, rc, max_neigh)
    ds = _dev()
    ds.set_atoms(x, y, z, box, origin, boundary)
    M, mx = ds.build_neighbor(rc, max_neigh)
    v, d, n = ds.fetch_neighbor()
    assert M == rv.shape[1] and mx == int(rn.max())
    assert np.array_equal(n, rn)
    assert np.array_equal(v, rv), "row order / membership differs from the reference"
    assert np.array_equal(d.view(np.int64), rd.view(np.int64)), "distances are not bit-identical"


def test_width_hint_across_frames_of_one_handle():
    """Frames of a trajectory reuse one handle: the second frame starts from the first frame's width
    (no sampling pass) and must still produce the reference's rows, also when its maximum grows or shrinks."""
    p, b = H.fcc(3.615, 12)
    rc = 3.3
    o, bnd = np.zeros(3), [1, 1, 1]
    ds = _dev()
    for sigma, seed in ((0.0, 0), (0.25, 1), (0.02, 2), (0.35, 3), (0.0, 4)):
        pos = H.rattle(p, sigma, seed) if sigma > 0 else p
        x, y, z = (np.ascontiguousarray(pos[:, k]) for k in range(3))
        rv, rd, rn = K.build_neighbor_auto(x, y, z, b, o, bnd, rc)
        ds.set_atoms(x, y, z, b, o, bnd)
        M, mx = ds.build_neighbor(rc)
        v, d, n = ds.fetch_neighbor()
        assert M == rv.shape[1] and np.array_equal(n, rn) and np.array_equal(v, rv)
        assert np.array_equal(d.view(np.int64), rd.view(np.int64))



def test_wrap_positions_bit_exact():
    """_neighbor.wrap_positions (neighbor.cpp:675): orthogonal with origin, triclinic, mixed boundaries, far images."""
    import mdapy_b200 as mp
    from mdapy_b200 import tool_function as tool

    rng = np.random.default_rng(41)
    p, b = H.fcc(3.615, 6)
    far = H.rattle(p, 0.3, 42) + rng.integers(-3, 4, p.shape) * np.diag(b)
    ps, bs = H.shear(far, b, xy=0.25, xz=-0.1, yz=0.2)
    for pos, box, origin, bnd in ((far, b, np.array([-3.0, 1.5, 0.25]), [1, 1, 1]), (far, b, np.zeros(3), [1, 0, 1]),
                                  (ps, bs, np.array([0.5, -2.0, 4.0]), [1, 1, 1]), (ps, bs, np.zeros(3), [0, 1, 1])):
        x, y, z = (np.ascontiguousarray(pos[:, k]) for k in range(3))
        rx, ry, rz = x.copy(), y.copy(), z.copy()
        K.wrap_positions(rx, ry, rz, box, origin, bnd)
        out = tool.wrap_pos(mp.Frame({"x": x, "y": y, "z": z}), mp.Box(box, bnd, origin))
        for got, ref in zip((out["x"], out["y"], out["z"]), (rx, ry, rz)):
            assert np.array_equal(np.asarray(got).view(np.int64), ref.view(np.int64))
    system = mp.System(pos=far, box=mp.Box(b))
    system.wrap_pos()
    assert float(np.asarray(system.data["x"]).min()) >= 0.0 and float(np.asarray(system.data["x"]).max()) < b[0, 0]
