"""Fused neighbour search + fixed-cutoff CNA (no neighbour list in HBM): labels must equal FixedCNA
(src/cna.cpp:429-506) on the reference's own list (src/neighbor.cpp:189-349), and the System must behave as if
the list had been built (it is, on first access)."""
import numpy as np
import pytest

import helpers as H
from oracle import checker as K

pytestmark = pytest.mark.gpu
O3 = np.zeros(3)


def _ref_labels(pos, box, boundary, rc):
    x, y, z = (np.ascontiguousarray(pos[:, k]) for k in range(3))
    rv, rd, rn = K.build_neighbor_auto(x, y, z, box, O3, boundary, rc)
    return K.fcna(x, y, z, box, O3, boundary, rv, rn, rc), (rv, rd, rn)


def _hcp(a=2.95, n=8):
    c = a * np.sqrt(8.0 / 3.0)
    basis = np.array([[0, 0, 0], [0.5, 0.5, 0], [0.5, 5.0 / 6.0, 0.5], [0, 1.0 / 3.0, 0.5]])
    cell = np.array([a, a * np.sqrt(3.0), c])
    ix, iy, iz = np.meshgrid(np.arange(n), np.arange(n), np.arange(n), indexing="ij")
    shift = np.stack([ix.ravel(), iy.ravel(), iz.ravel()], axis=1).astype(float)
    pos = ((basis[None] + shift[:, None]) * cell).reshape(-1, 3)
    return np.ascontiguousarray(pos), np.diag(cell * n)


def _cases():
    out = []
    p, b = H.fcc(3.615, 12)
    rc = 3.615 * 0.8536
    out.append(("fcc", p, b, [1, 1, 1], rc))
    out.append(("fcc_rattled", H.rattle(p, 0.05, 0), b, [1, 1, 1], rc))
    out.append(("fcc_hot", H.rattle(p, 0.2, 1), b, [1, 1, 1], rc))
    out.append(("fcc_very_hot_unwrapped", H.rattle(p, 0.45, 2), b, [1, 1, 1], rc))
    out.append(("fcc_slab", H.rattle(p, 0.08, 3), b, [1, 1, 0], rc))
    out.append(("fcc_wire", H.rattle(p, 0.08, 4), b, [1, 0, 0], rc))
    out.append(("fcc_open", H.rattle(p, 0.08, 5), b, [0, 0, 0], rc))
    out.append(("fcc_far_images", H.rattle(p, 0.05, 9) + np.array([3, -2, 5]) * np.diag(b), b, [1, 1, 1], rc))
    p2, b2 = H.bcc(2.8665, 14)
    out.append(("bcc", p2, b2, [1, 1, 1], 2.8665 * 1.207))
    out.append(("bcc_rattled", H.rattle(p2, 0.04, 6), b2, [1, 1, 1], 2.8665 * 1.207))
    p3, b3 = _hcp(n=10)
    out.append(("hcp", p3, b3, [1, 1, 1], 2.95 * 1.207))
    out.append(("hcp_rattled", H.rattle(p3, 0.05, 7), b3, [1, 1, 1], 2.95 * 1.207))
    # a stacking fault: fcc + hcp layers, plus vacancies
    pv = np.delete(H.rattle(p, 0.03, 8), np.arange(0, p.shape[0], 97), axis=0)
    out.append(("fcc_vacancies", pv, b, [1, 1, 1], rc))
    g, bg = H.random_gas(6000, 40.0, 7)
    out.append(("gas", g, bg, [1, 1, 1], 3.2))
    ps, bs = H.shear(H.rattle(p, 0.08, 11), b, xy=0.2, xz=0.1, yz=-0.15)
    out.append(("fcc_triclinic", ps, bs, [1, 1, 1], rc))
    ps, bs = H.shear(H.rattle(p2, 0.05, 12), b2, xy=0.5, xz=-0.2, yz=0.3)
    out.append(("bcc_tilted_open_y", ps, bs, [1, 0, 1], 2.8665 * 1.207))
    # exact ties: rc exactly at a shell (every neighbour test and bond test lands in the guard band)
    out.append(("fcc_rc_on_second_shell", p, b, [1, 1, 1], 3.615))
    return out


CASES = _cases()


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_fused_labels_equal_fixed_cna_on_the_reference_list(case):
    from mdapy_b200.device import DeviceSystem

    name, pos, box, boundary, rc = case
    ref, _ = _ref_labels(pos, box, boundary, rc)
    ds = DeviceSystem(0)
    ds.set_atoms(pos[:, 0].copy(), pos[:, 1].copy(), pos[:, 2].copy(), box, O3, boundary)
    lab, used = ds.fused_cna(rc)
    assert used, "orthogonal frame with >= 7 cells per periodic axis must take the fused kernel"
    assert np.array_equal(lab, ref), f"{name}: {np.bincount(lab, minlength=5)} vs {np.bincount(ref, minlength=5)}"
    # the list path on the same handle gives the same labels (and is untouched by the fused call)
    ds.build_neighbor(rc, None)
    assert np.array_equal(ds.fcna(rc), ref)


def test_not_eligible_frames_report_unused():
    """Fewer than 7 cells along a periodic axis: no unambiguous nearest image inside a tile -> the list path."""
    from mdapy_b200.device import DeviceSystem

    p, b = H.fcc(3.615, 6)
    ps, bs = H.shear(H.rattle(p, 0.05, 5), b, xy=0.2, xz=0.1, yz=-0.15)
    ds = DeviceSystem(0)
    ds.set_atoms(ps[:, 0].copy(), ps[:, 1].copy(), ps[:, 2].copy(), bs, O3, [1, 1, 1])
    lab, used = ds.fused_cna(3.4)
    assert not used and lab is None


def test_system_builds_the_list_lazily_and_matches_the_reference_state():
    import mdapy_b200 as mp
    from mdapy_b200 import _lib

    p, b = H.fcc(3.615, 12)
    pos = H.rattle(p, 0.05, 0)
    rc = 3.615 * 0.8536
    ref, (rv, rd, rn) = _ref_labels(pos, b, [1, 1, 1], rc)
    system = mp.System(pos=pos, box=b)
    system.cal_common_neighbor_analysis(rc)
    assert np.array_equal(np.asarray(system.data["cna"]), ref)
    assert system._pending_rc == rc and not system._has_list      # nothing materialised yet
    assert system.rc == rc                                        # first access builds the list
    assert system._pending_rc is None and system._has_list
    assert np.array_equal(system.verlet_list, rv) and np.array_equal(system.neighbor_number, rn)
    assert np.array_equal(system.distance_list.view(np.int64), rd.view(np.int64))
    # a consumer that reuses the cached list (system.py:1986-2003) sees it, too
    system2 = mp.System(pos=pos, box=b)
    system2.cal_common_neighbor_analysis(rc)
    system2.cal_centro_symmetry_parameter(12)
    x, y, z = (np.ascontiguousarray(pos[:, k]) for k in range(3))
    K.sort_verlet_by_distance(rv, rd, 12)
    ref_csp = K.csp(x, y, z, b, O3, [1, 1, 1], rv, 12)
    assert np.array_equal(np.asarray(system2.data["csp"]).view(np.int64), ref_csp.view(np.int64))
    # larger cut-off afterwards replaces the pending list; a k-nearest build keeps rc like the reference
    system3 = mp.System(pos=pos, box=b)
    system3.cal_common_neighbor_analysis(rc)
    system3.build_nearest_neighbor(12)
    assert system3.rc == rc and system3.verlet_list.shape[1] == 12
    assert _lib.lib().mdb_launch_count() > 0
