"""Device group (csrc/group.cu, System(devices=[...])): ONE unpartitioned host frame sharded over several GPUs
by one process.  Labels must equal FixedCNA (src/cna.cpp:429-506) on the reference's own list
(src/neighbor.cpp:189-349), in the original atom order, whatever the member count.  A GPU may be listed more
than once, which runs the whole routing path (chunk upload, count, peer-store push, slab compute, label push)
on a single-GPU box; the real multi-GPU cases run when the box has the devices."""
import numpy as np
import pytest

import helpers as H
from oracle import checker as K

pytestmark = pytest.mark.gpu
O3 = np.zeros(3)


def _ref_labels(pos, box, boundary, rc):
    x, y, z = (np.ascontiguousarray(pos[:, k]) for k in range(3))
    rv, rd, rn = K.build_neighbor_auto(x, y, z, box, O3, np.asarray(boundary, np.int32), rc)
    return K.fcna(x, y, z, box, O3, np.asarray(boundary, np.int32), rv, rn, rc)


def _group_labels(devices, pos, box, boundary, rc, origin=O3):
    from mdapy_b200.device import DeviceGroup

    g = DeviceGroup(devices)
    x, y, z = (np.ascontiguousarray(pos[:, k]) for k in range(3))
    g.set_atoms(x, y, z, box, origin, np.asarray(boundary, np.int32))
    lab = np.array(g.fused_cna(rc))
    info = (g.members_used, g.member_atoms(), g.last_times())
    g.close()
    return lab, info


def _cases():
    out = []
    rc = 3.615 * 0.8536
    p, b = H.fcc(3.615, 14)                       # 16 x-planes
    out.append(("fcc", p, b, [1, 1, 1], rc))
    out.append(("fcc_rattled", H.rattle(p, 0.05, 0), b, [1, 1, 1], rc))
    out.append(("fcc_hot", H.rattle(p, 0.2, 1), b, [1, 1, 1], rc))
    out.append(("fcc_very_hot_unwrapped", H.rattle(p, 0.45, 2), b, [1, 1, 1], rc))
    out.append(("fcc_slab_open_z", H.rattle(p, 0.08, 3), b, [1, 1, 0], rc))
    out.append(("fcc_open_x", H.rattle(p, 0.08, 4), b, [0, 1, 1], rc))
    out.append(("fcc_open", H.rattle(p, 0.08, 5), b, [0, 0, 0], rc))
    out.append(("fcc_far_images", H.rattle(p, 0.05, 9) + np.array([3, -2, 5]) * np.diag(b), b, [1, 1, 1], rc))
    p2, b2 = H.bcc(2.8665, 20)
    out.append(("bcc_rattled", H.rattle(p2, 0.04, 6), b2, [1, 1, 1], 2.8665 * 1.207))
    # shuffled atom order: a chunk of the input is no longer a slab of space, every member routes to every other
    rng = np.random.default_rng(3)
    out.append(("fcc_shuffled", H.rattle(p, 0.1, 10)[rng.permutation(p.shape[0])], b, [1, 1, 1], rc))
    # N not a multiple of anything
    out.append(("fcc_vacancies", np.delete(H.rattle(p, 0.03, 8), np.arange(0, p.shape[0], 97), axis=0), b, [1, 1, 1], rc))
    g, bg = H.random_gas(11000, 50.0, 7)
    out.append(("gas", g, bg, [1, 1, 1], 3.2))
    return out


CASES = _cases()


@pytest.mark.parametrize("members", [2, 3, 5])
@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_group_labels_equal_reference(case, members):
    _, pos, box, boundary, rc = case
    ref = _ref_labels(pos, box, boundary, rc)
    lab, (used, atoms, _) = _group_labels([0] * members, pos, box, boundary, rc)
    assert used == members, "frame should have been sharded"
    assert sum(a for a, _ in atoms) == pos.shape[0], "every atom is owned exactly once"
    assert all(t > a for a, t in atoms), "every slab holds ghosts"
    assert np.array_equal(lab, ref)


def test_group_origin_shift():
    p, b = H.fcc(3.615, 14)
    rc = 3.615 * 0.8536
    pos = H.rattle(p, 0.1, 21)
    origin = np.array([-7.5, 3.25, 100.0])
    x, y, z = (np.ascontiguousarray(pos[:, k] + origin[k]) for k in range(3))
    rv, rd, rn = K.build_neighbor_auto(x, y, z, b, origin, np.array([1, 1, 1], np.int32), rc)
    ref = K.fcna(x, y, z, b, origin, np.array([1, 1, 1], np.int32), rv, rn, rc)
    lab, (used, _, _) = _group_labels([0, 0, 0], np.stack([x, y, z], 1), b, [1, 1, 1], rc, origin=origin)
    assert used == 3 and np.array_equal(lab, ref)


def test_group_gathers_what_it_cannot_shard():
    rc = 3.615 * 0.8536
    # 9 x-planes: enough for 3 members, not for 4 -> gathered on the first member, same labels
    p, b = H.fcc(3.615, 8)
    pos = H.rattle(p, 0.1, 13)
    ref = _ref_labels(pos, b, [1, 1, 1], rc)
    lab3, (used3, _, _) = _group_labels([0, 0, 0], pos, b, [1, 1, 1], rc)
    lab4, (used4, _, _) = _group_labels([0, 0, 0, 0], pos, b, [1, 1, 1], rc)
    assert (used3, used4) == (3, 1)
    assert np.array_equal(lab3, ref) and np.array_equal(lab4, ref)
    # a frame the fused kernel declines inside its slab (more atoms than members would ever get is fine; a
    # tiny frame is not): 3 atoms on 2 members -> one member would be empty -> gathered
    tiny = np.array([[0.5, 0.5, 0.5], [2.0, 0.5, 0.5], [0.5, 2.0, 0.5]])
    bt = np.diag([30.0, 30.0, 30.0])
    lab, (used, _, _) = _group_labels([0, 0], tiny, bt, [1, 1, 1], 3.0)
    assert used == 1 and np.array_equal(lab, _ref_labels(tiny, bt, [1, 1, 1], 3.0))


def test_group_triclinic_frame():
    p, b = H.fcc(3.615, 14)
    rc = 3.615 * 0.8536
    ps, bs = H.shear(H.rattle(p, 0.08, 11), b, xy=0.2, xz=0.1, yz=-0.15)
    ref = _ref_labels(ps, bs, [1, 1, 1], rc)
    lab, (used, _, _) = _group_labels([0, 0, 0], ps, bs, [1, 1, 1], rc)
    assert used in (1, 3)          # sharded when the fused kernel takes triclinic slabs, gathered otherwise
    assert np.array_equal(lab, ref)


def test_group_single_member_and_reuse():
    from mdapy_b200.device import DeviceGroup

    rc = 3.615 * 0.8536
    p, b = H.fcc(3.615, 12)
    g = DeviceGroup([0])
    for seed in (0, 1):                      # one group, two frames
        pos = H.rattle(p, 0.1, seed)
        x, y, z = (np.ascontiguousarray(pos[:, k]) for k in range(3))
        g.set_atoms(x, y, z, b, O3, np.array([1, 1, 1], np.int32))
        assert np.array_equal(np.array(g.fused_cna(rc)), _ref_labels(pos, b, [1, 1, 1], rc))
        assert g.members_used == 1
    g.close()
    g = DeviceGroup([0, 0, 0])
    for n, seed in ((14, 2), (16, 3), (14, 4)):     # frames of different sizes through the same buffers
        p, b = H.fcc(3.615, n)
        pos = H.rattle(p, 0.15, seed)
        x, y, z = (np.ascontiguousarray(pos[:, k]) for k in range(3))
        g.set_atoms(x, y, z, b, O3, np.array([1, 1, 1], np.int32))
        assert np.array_equal(np.array(g.fused_cna(rc)), _ref_labels(pos, b, [1, 1, 1], rc))
        assert g.members_used == 3
    g.close()


def test_group_errors():
    from mdapy_b200.device import DeviceGroup

    with pytest.raises((RuntimeError, ValueError)):
        DeviceGroup([9999])
    with pytest.raises(ValueError):
        DeviceGroup([])
    g = DeviceGroup([0, 0])
    with pytest.raises(RuntimeError):
        g.fused_cna(3.0)                     # nothing uploaded
    g.close()


def test_system_devices_keyword():
    import mdapy_b200 as mp

    rc = 3.615 * 0.8536
    p, b = H.fcc(3.615, 14)
    pos = H.rattle(p, 0.12, 5)
    ref = _ref_labels(pos, b, [1, 1, 1], rc)
    single = mp.System(pos=pos, box=mp.Box(b))
    single.cal_common_neighbor_analysis(rc)
    multi = mp.System(pos=pos, box=mp.Box(b), devices=[0, 0, 0])
    multi.cal_common_neighbor_analysis(rc)
    assert np.array_equal(np.asarray(multi.data["cna"]), ref)
    assert np.array_equal(np.asarray(single.data["cna"]), ref)
    assert multi._group.members_used == 3
    # the state afterwards is the reference's: the list exists (built on first access, on devices[0])
    assert multi.rc == rc
    assert np.array_equal(multi.neighbor_number, single.neighbor_number)
    assert np.array_equal(multi.verlet_list, single.verlet_list)
    # a second descriptor on the same System runs on devices[0]
    multi.cal_centro_symmetry_parameter(12)
    single.cal_centro_symmetry_parameter(12)
    assert np.array_equal(np.asarray(multi.data["csp"]), np.asarray(single.data["csp"]))


def _n_gpus():
    from mdapy_b200 import _lib as L
    import ctypes as C

    n = C.c_int(0)
    L.check(L.lib().mdb_device_count(C.byref(n)))
    return n.value


def test_group_real_devices():
    n = _n_gpus()
    if n < 2:
        pytest.skip("needs at least two GPUs")
    rc = 3.615 * 0.8536
    p, b = H.fcc(3.615, 40)                     # 256 k atoms, 46 planes
    rng = np.random.default_rng(0)
    pos = H.rattle(p, 0.12, 17)
    for devs in ([0, 1], list(range(n)), list(range(n)) * 2):
        if 3 * len(devs) > 46:
            continue
        ref = _ref_labels(pos, b, [1, 1, 1], rc)
        lab, (used, _, _) = _group_labels(devs, pos, b, [1, 1, 1], rc)
        assert used == len(devs) and np.array_equal(lab, ref)
        shuffled = pos[rng.permutation(pos.shape[0])]
        lab, (used, _, _) = _group_labels(devs, shuffled, b, [1, 1, 1], rc)
        assert used == len(devs) and np.array_equal(lab, _ref_labels(shuffled, b, [1, 1, 1], rc))


def test_second_device_runs_every_kernel_family():
    """Per-device state (dynamic shared memory attributes, constant tables) must follow the device of the
    handle, not the first device the process touched."""
    if _n_gpus() < 2:
        pytest.skip("needs at least two GPUs")
    import mdapy_b200 as mp

    p, b = H.fcc(3.615, 10)
    pos = H.rattle(p, 0.1, 3)
    cols = {}
    for dev in (0, 1):
        s = mp.System(pos=pos, box=mp.Box(b), device=dev)
        s.cal_common_neighbor_analysis(3.615 * 0.8536)
        s.cal_steinhardt_bond_orientation([4, 6], rc=3.615 * 0.8536, average=True)
        s.cal_polyhedral_template_matching(return_rmsd=True)
        s.cal_centro_symmetry_parameter(12)
        s.cal_ackland_jones_analysis()
        cols[dev] = {k: np.asarray(s.data[k]) for k in s.data.columns if k not in ("x", "y", "z")}
        v = s.verlet_list
        cols[dev]["verlet"] = np.asarray(v)
    assert cols[0].keys() == cols[1].keys() and len(cols[0]) >= 6
    for k in cols[0]:
        assert np.array_equal(cols[0][k], cols[1][k]), k
