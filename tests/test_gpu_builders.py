"""SURVEY.md 8f.2 -- builders on the device against the reference's own C++ (oracle/_ref): repeat_cell
(repeat_cell.cpp:19), transform_and_filter (polycrystal.cpp:21), filter_overlap_atom (neighbor.cpp:390).
Bit-exact bar: coordinates by bit pattern, kept sets equal."""
import numpy as np
import pytest

import helpers as H
from oracle import checker as K

pytestmark = pytest.mark.gpu
O3 = np.zeros(3)


def _bits(a):
    return np.ascontiguousarray(a).view(np.int64)


@pytest.mark.parametrize("case", ["fcc", "triclinic_basis"])
def test_repeat_cell_bit_exact(case):
    from mdapy_b200 import builders as B

    if case == "fcc":
        box = 3.615 * np.eye(3)
        pos = H.FCC @ box
        reps = (7, 5, 9)
    else:
        box = np.array([[3.1, 0.0, 0.0], [0.7, 2.9, 0.0], [-0.3, 0.4, 3.3]])
        pos = np.random.default_rng(0).random((5, 3)) @ box
        reps = (4, 6, 3)
    ref = K.repeat_cell(box, pos, *reps)
    got = B.repeat_cell(box, pos, *reps)
    assert got.shape == ref.shape and np.array_equal(_bits(got), _bits(ref))


def test_device_lattice_equals_build_crystal():
    from mdapy_b200 import builders as B

    ds = B.device_lattice("fcc", 4.05, 9, 8, 7)
    x, y, z = B.fetch_positions(ds)
    ref, box = H.lattice(H.FCC, 4.05, 9, 8, 7)
    assert np.array_equal(_bits(x), _bits(ref[:, 0])) and np.array_equal(_bits(y), _bits(ref[:, 1]))
    assert np.array_equal(_bits(z), _bits(ref[:, 2]))
    # and the frame is usable as is: perfect FCC, every atom labelled fcc
    lab, used = ds.fused_cna(4.05 * 0.8536)
    assert used and np.all(lab == 1)


def test_transform_and_filter_bit_exact():
    from mdapy_b200 import builders as B

    rng = np.random.default_rng(3)
    p, b = H.fcc(4.05, 14)
    Lx = np.diag(b)
    seeds = rng.random((6, 3)) * Lx
    theta = rng.uniform(-np.pi, np.pi, 3)
    cx, sx = np.cos(theta[0]), np.sin(theta[0])
    cy, sy = np.cos(theta[1]), np.sin(theta[1])
    cz, sz = np.cos(theta[2]), np.sin(theta[2])
    R = (np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]]) @ np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
         @ np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]]))
    for g in (0, 3):
        planes = B.bisector_planes(seeds, g, Lx)
        x, y, z = (np.ascontiguousarray(p[:, k]) for k in range(3))
        ref = K.transform_and_filter(x, y, z, R, Lx / 2, seeds[g], planes)
        got = B.transform_and_filter(x, y, z, R, Lx / 2, seeds[g], planes)
        assert ref.shape[0] > 50
        assert got.shape == ref.shape and np.array_equal(_bits(got), _bits(ref))
    # no planes: everything is kept, only transformed
    got = B.transform_and_filter(x, y, z, R, Lx / 2, seeds[0], np.zeros((0, 4)))
    ref = K.transform_and_filter(x, y, z, R, Lx / 2, seeds[0], np.zeros((0, 4)))
    assert np.array_equal(_bits(got), _bits(ref))


@pytest.mark.parametrize("case", ["hot_fcc", "gas", "triclinic", "open"])
def test_filter_overlap_atom_equal(case):
    from mdapy_b200 import builders as B

    boundary = [1, 1, 1]
    if case == "hot_fcc":
        p, b = H.fcc(3.615, 8)
        pos, rc = H.rattle(p, 0.35, 1), 2.2
    elif case == "gas":
        pos, b = H.random_gas(5000, 30.0, 2)
        rc = 1.6
    elif case == "triclinic":
        p, b0 = H.fcc(3.615, 8)
        pos, b = H.shear(H.rattle(p, 0.3, 4), b0, xy=0.2, xz=0.1, yz=-0.1)
        rc = 2.1
    else:
        p, b = H.fcc(3.615, 8)
        pos, rc, boundary = H.rattle(p, 0.5, 5), 2.2, [0, 1, 0]
    x, y, z = (np.ascontiguousarray(pos[:, k]) for k in range(3))
    ref = K.filter_overlap_atom(x, y, z, b, O3, boundary, rc)
    got = B.filter_overlap_atom(x, y, z, b, O3, boundary, rc)
    assert 0 < ref.sum() < ref.size
    assert np.array_equal(got, ref)


def test_create_polycrystal_matches_the_reference_recipe():
    """The whole builder (create_polycrystal.py:684-840) against the same recipe driven through the reference's
    C++ helpers: identical atoms, bit for bit."""
    from mdapy_b200.create_polycrystal import CreatePolycrystal, _rodrigues

    a, Lbox, G = 4.05, 60.0, 5
    poly = CreatePolycrystal("fcc", a, Lbox, G, randomseed=7)
    system = poly.compute()
    planes, radius = poly._cells()
    reps = int(np.ceil(radius.max() / a))
    block = K.repeat_cell(a * np.eye(3), H.FCC @ (a * np.eye(3)), reps, reps, reps)
    centre = poly.block_centre        # the one input that is a floating-point mean (summation order is the caller's)
    assert np.allclose(centre, block.mean(axis=0), rtol=1e-12)
    x, y, z = (np.ascontiguousarray(block[:, k]) for k in range(3))
    parts = []
    for g in range(G):
        th = poly.theta_list[g]
        R = _rodrigues(th[0], (1.0, 0, 0)) @ _rodrigues(th[1], (0, 1.0, 0)) @ _rodrigues(th[2], (0, 0, 1.0))
        parts.append(K.transform_and_filter(x, y, z, R, centre, poly.seed_position[g], planes[g]))
    pos = np.concatenate(parts)
    box = np.eye(3) * Lbox
    keep = K.filter_overlap_atom(pos[:, 0].copy(), pos[:, 1].copy(), pos[:, 2].copy(), box, O3, [1, 1, 1], 2.0)
    pos = pos[keep]
    wx, wy, wz = pos[:, 0].copy(), pos[:, 1].copy(), pos[:, 2].copy()
    K.wrap_positions(wx, wy, wz, box, O3, [1, 1, 1])   # in place
    got = np.stack([np.asarray(system.data[c]) for c in "xyz"], axis=1)
    assert got.shape[0] == wx.shape[0] and 0.9 < got.shape[0] / (4 * (Lbox / a) ** 3) < 1.0
    assert np.array_equal(_bits(got[:, 0]), _bits(wx)) and np.array_equal(_bits(got[:, 1]), _bits(wy))
    assert np.array_equal(_bits(got[:, 2]), _bits(wz))
    # the bulk of every grain is fcc
    system.cal_common_neighbor_analysis(a * 0.8536)
    assert (np.asarray(system.data["cna"]) == 1).mean() > 0.6
