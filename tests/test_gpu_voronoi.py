"""Voronoi cells on the GPU (csrc/voronoi.cu) against the reference's src/voronoi.cpp + vendored voro++:
the upstream fixtures (tests/test_voronoi.py of the reference: volume, cavity radius, coordination, atol 1e-6),
reference-run vectors committed under tests/golden/voronoi.npz, and live runs of the compiled reference.
Face ORDER inside a row is voro++'s vertex-table order in the reference and insertion order here: rows are compared
as sets (sorted by neighbour id).  Tolerances: face counts and neighbour ids exact, volume / area / radius 1e-9
relative (different but equally valid floating-point constructions of the same polyhedron)."""
from pathlib import Path

import numpy as np
import pytest

import helpers as H
from oracle import checker as K

pytestmark = pytest.mark.gpu
O3 = np.zeros(3)
GOLD_DIR = Path(__file__).resolve().parent / "golden"
GOLD = np.load(GOLD_DIR / "voronoi.npz")
RTOL = 1e-9


def _canon(verlet, dist, area):
    key = np.where(verlet < 0, np.iinfo(np.int32).max, verlet)
    order = np.lexsort((area, key), axis=1)      # by neighbour id, then face area (a thin box meets one atom twice)
    return [np.take_along_axis(a, order, axis=1) for a in (verlet, dist, area)]


def _device(pos, box, boundary, origin=O3):
    """The device handle mdapy_b200.voronoi.Voronoi would use (triclinic: open axes tripled, all periodic)."""
    import mdapy_b200 as mp
    from mdapy_b200.voronoi import Voronoi

    b = mp.Box(np.vstack([np.asarray(box, float)[:3], np.asarray(origin, float).reshape(1, 3)]),
               boundary=np.asarray(boundary, np.int32))
    v = Voronoi(b, {"x": pos[:, 0], "y": pos[:, 1], "z": pos[:, 2]})
    return v._device_for(b, v.data, reuse=False)


def _tags(name):
    tags = {"none": (-1.0, -1.0)}
    if name == "rattled":
        tags.update({"abs": (0.6, -1.0), "rel": (-1.0, 0.02), "both": (0.3, 0.05)})
    return tags


@pytest.mark.parametrize("name", [str(n) for n in GOLD["fixture_names"]])
def test_upstream_fixture(name):
    """The reference's own test (tests/test_voronoi.py:13-36) on all 15 fixtures (3 of them triclinic)."""
    d = np.load(GOLD_DIR / f"sa_{name}.npz")
    box = np.asarray(d["box"], float)
    origin = box[3] if box.shape[0] == 4 else O3
    ds = _device(d["pos"], box, d["boundary"], origin)
    vol, nn, rad = ds.voronoi_volume()
    assert np.array_equal(nn, GOLD[f"{name}__voronoi_coord"])
    assert np.allclose(vol, GOLD[f"{name}__voronoi_volume"], atol=1e-6)
    assert np.allclose(rad * 0.5, GOLD[f"{name}__voronoi_cavity_radius"], atol=1e-6)


@pytest.mark.parametrize("name", [str(n) for n in GOLD["run_names"]])
def test_reference_run_vectors(name):
    pos, box, bd = GOLD[f"run_{name}__pos"], GOLD[f"run_{name}__box"], GOLD[f"run_{name}__boundary"]
    ds = _device(pos, box, bd)
    vol, nn, rad = ds.voronoi_volume()
    assert np.array_equal(nn, GOLD[f"run_{name}__faces"])
    assert np.allclose(vol, GOLD[f"run_{name}__volume"], rtol=RTOL, atol=0)
    assert np.allclose(rad, GOLD[f"run_{name}__radius"], rtol=RTOL, atol=0)
    if all(bd):
        assert abs(vol.sum() - abs(np.linalg.det(box))) < 1e-9 * abs(np.linalg.det(box))
    for tag, (at, rt) in _tags(name).items():
        v, dd, ar, n2 = ds.voronoi_neighbor(at, rt)
        v, dd, ar = _canon(np.array(v), np.array(dd), np.array(ar))
        assert np.array_equal(n2, GOLD[f"run_{name}__{tag}_nn"]), tag
        assert np.array_equal(v, GOLD[f"run_{name}__{tag}_verlet"]), tag
        assert np.allclose(ar, GOLD[f"run_{name}__{tag}_area"], rtol=1e-7, atol=1e-9), tag
        assert np.allclose(dd, GOLD[f"run_{name}__{tag}_dist"], rtol=1e-14, atol=0), tag     # minimum-image distances


def _live_cases():
    out = []
    p, b = H.fcc(3.615, 8)
    out.append(("fcc_rattled", H.rattle(p, 0.1, 0), b, [1, 1, 1]))
    out.append(("fcc_hot_unwrapped", H.rattle(p, 0.4, 1) + np.array([2, -1, 3]) * np.diag(b), b, [1, 1, 1]))
    p2, b2 = H.bcc(2.8665, 9)
    out.append(("bcc_rattled_open_y", np.clip(H.rattle(p2, 0.05, 2), 1e-3, np.diag(b2) - 1e-3), b2, [1, 0, 1]))
    g, bg = H.random_gas(3000, 40.0, 3)
    out.append(("gas", g, bg, [1, 1, 1]))
    out.append(("gas_open", g, bg, [0, 0, 0]))
    sc, bsc = H.lattice(np.zeros((1, 3)), 2.6, 8, 8, 8)
    out.append(("perfect_sc", sc, bsc, [1, 1, 1]))                 # every second neighbour touches an edge or a vertex
    out.append(("perfect_fcc_open_x", p + 0.2, b, [0, 1, 1]))
    dia, bdia = H.diamond(3.567, 5)
    out.append(("diamond_noise_1e-7", H.rattle(dia, 1e-7, 3), bdia, [1, 1, 1]))
    rng = np.random.default_rng(4)
    thin = rng.random((400, 3)) * [60.0, 5.0, 7.0]
    out.append(("thin_box_own_images", thin, np.diag([60.0, 5.0, 7.0]), [1, 1, 1]))
    ps, bs = H.shear(H.rattle(p, 0.1, 7), b, xy=0.2, xz=0.1, yz=-0.15)
    out.append(("fcc_rattled_triclinic", ps, bs, [1, 1, 1]))
    ps, bs = H.shear(H.rattle(p2, 0.05, 8), b2, xy=0.5, xz=-0.2, yz=0.3)
    out.append(("bcc_tilted_open_y", ps, bs, [1, 0, 1]))
    gg, bgg = H.random_gas(600, 12.0, 9)
    gs, bgs = H.shear(gg, bgg, xy=0.4, xz=0.3, yz=0.35)
    out.append(("gas_small_triclinic", gs, bgs, [1, 1, 1]))
    return out


LIVE = _live_cases()


@pytest.mark.parametrize("case", LIVE, ids=[c[0] for c in LIVE])
def test_against_compiled_reference(case):
    if K.KIND != "reference" or not hasattr(K, "voronoi_volume"):
        pytest.skip("needs oracle/_ref/ref_voronoi.so (the reference's voro++ compiled in the dev container)")
    _, pos, box, bd = case
    x, y, z = (np.ascontiguousarray(pos[:, k]) for k in range(3))
    bd = np.asarray(bd, np.int32)
    tri = np.abs(box - np.diag(np.diag(box))).max() > 1e-10
    rvol, rnn, rrad = (K.voronoi_volume_tri if tri else K.voronoi_volume)(x, y, z, box, O3, bd)
    ds = _device(pos, box, bd)
    vol, nn, rad = ds.voronoi_volume()
    assert np.array_equal(nn, rnn)
    assert np.allclose(vol, rvol, rtol=RTOL, atol=0) and np.allclose(rad, rrad, rtol=RTOL, atol=0)
    rv, rd, ra, rn = (K.voronoi_neighbor_tri if tri else K.voronoi_neighbor)(x, y, z, box, O3, bd, -1.0, 0.01)
    v, dd, ar, n2 = ds.voronoi_neighbor(-1.0, 0.01)
    rv, rd, ra = _canon(rv, rd, ra)
    v, dd, ar = _canon(np.array(v), np.array(dd), np.array(ar))
    assert v.shape == rv.shape and np.array_equal(n2, rn) and np.array_equal(v, rv)
    assert np.allclose(dd, rd, rtol=1e-14, atol=0) and np.allclose(ar, ra, rtol=1e-7, atol=1e-9)


def test_system_api_and_steinhardt_with_voronoi():
    import mdapy_b200 as mp

    p, b = H.fcc(3.615, 6)
    pos = H.rattle(p, 0.1, 5)
    s = mp.System(pos=pos, box=mp.Box(b))
    s.cal_voronoi_volume()
    assert abs(np.asarray(s.data["volume"]).sum() - np.prod(np.diag(b))) < 1e-8
    assert np.asarray(s.data["neighbor_number"]).min() >= 12
    s.build_voronoi_neighbor(r_face_area_threshold=0.01)
    assert s.voro_verlet_list.shape == s.voro_face_area.shape == s.voro_distance_list.shape
    s.cal_steinhardt_bond_orientation([4, 6], use_voronoi=True, use_weight=True, r_face_area_threshold=0.01)
    q6 = np.asarray(s.data["ql6"])
    # the same q_l from the reference's kernel on the same Voronoi rows (sbo.cpp:288-576; row order only permutes
    # the summation, so agreement is to rounding)
    x, y, z = (np.ascontiguousarray(pos[:, k]) for k in range(3))
    qn, _, _ = K.get_sq(x, y, z, b, O3, np.array([1, 1, 1], np.int32), np.array(s.voro_verlet_list),
                        np.array(s.voro_distance_list), np.array(s.voro_neighbor_number), [4, 6], nnn=0, rc=-1.0,
                        use_voronoi=True, weight=np.array(s.voro_face_area))
    assert np.allclose(q6, qn[:, 1], rtol=1e-12, atol=1e-14)
    assert 0.3 < q6.mean() < 0.6


def test_small_periodic_frame_is_replicated_like_the_reference():
    import mdapy_b200 as mp

    p, b = H.fcc(3.615, 2)            # 32 atoms < 50: replicated to 256 (voronoi.py:118-137)
    s = mp.System(pos=H.rattle(p, 0.05, 6), box=mp.Box(b))
    s.build_voronoi_neighbor()
    assert s.voro_verlet_list.shape[0] == 256 and hasattr(s, "_enlarge_data")
    s.cal_voronoi_volume()            # volumes: no replication, the cells meet their own images
    assert abs(np.asarray(s.data["volume"]).sum() - np.prod(np.diag(b))) < 1e-9


def test_triclinic_with_open_axes_needs_the_python_side():
    """The library builds triclinic cells for fully periodic boxes; mdapy_b200.voronoi triples the open axes first
    (as the reference's Python side does), a raw handle with an open triclinic axis is refused."""
    from mdapy_b200.device import DeviceSystem

    p, b = H.fcc(3.615, 6)
    ps, bs = H.shear(p, b, xy=0.2, xz=0.0, yz=0.0)
    ds = DeviceSystem(0)
    ds.set_atoms(*(np.ascontiguousarray(ps[:, k]) for k in range(3)), bs, O3, np.array([1, 0, 1], np.int32))
    with pytest.raises(ValueError):
        ds.voronoi_volume()


def test_steinhardt_on_voronoi_rows_end_to_end():
    """q_l / w_l-hat with Voronoi neighbours, face-area weights, an absolute area threshold and neighbour averaging:
    bit-identical to the reference kernel on the same rows, and equal to rounding (row order only permutes the sums)
    to the reference end to end (voro++ rows -> get_sq)."""
    import mdapy_b200 as mp

    if K.KIND != "reference":
        pytest.skip("needs oracle/_ref")
    p, b = H.bcc(2.8665, 8)
    pos = H.rattle(p, 0.08, 7)
    PB = np.array([1, 1, 1], np.int32)
    s = mp.System(pos=pos, box=mp.Box(b))
    sbo = s.cal_steinhardt_bond_orientation([4, 6, 8], use_voronoi=True, use_weight=True, average=True, wlhat=True,
                                            a_face_area_threshold=0.3)
    x, y, z = (np.ascontiguousarray(pos[:, k]) for k in range(3))
    qn, _, _ = K.get_sq(x, y, z, b, O3, PB, np.array(s.voro_verlet_list), np.array(s.voro_distance_list),
                        np.array(s.voro_neighbor_number), [4, 6, 8], average=True, wlhat=True, use_voronoi=True,
                        weight=np.array(s.voro_face_area))
    assert np.array_equal(sbo.qnarray, qn)
    rv, rd, ra, rn = K.voronoi_neighbor(x, y, z, b, O3, PB, 0.3, -1.0)
    qr, _, _ = K.get_sq(x, y, z, b, O3, PB, rv, rd, rn, [4, 6, 8], average=True, wlhat=True, use_voronoi=True, weight=ra)
    assert np.allclose(sbo.qnarray, qr, rtol=0, atol=1e-13)
    # fewer than 50 atoms: replicated like the reference, columns come back for the original atoms
    p2, b2 = H.fcc(3.615, 2)
    t = mp.System(pos=H.rattle(p2, 0.05, 6), box=mp.Box(b2))
    t.cal_steinhardt_bond_orientation([4, 6], use_voronoi=True, average=True, identify_liquid=True, wl=True)
    assert np.asarray(t.data["ql6"]).shape == (32,) and np.asarray(t.data["solidliquid"]).shape == (32,)
