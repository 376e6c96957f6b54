"""GPU parity of polyhedral template matching against the reference's own C++ (oracle/_ref) and against
the committed reference-run vectors (tests/golden/sa_*.npz: ref_ptm_output).

Bar: structure type and alloy ordering bit-exact; rmsd, interatomic distance and the orientation
quaternion (up to the sign of q) within 1e-6 relative -- observed ~1e-14.  Exception, documented in
DESIGN.md: on unperturbed lattices WITH a defect the Voronoi solid angles of symmetric neighbours tie
exactly and the reference's own order among them is floating-point noise; those fixtures compare types at
the default threshold only."""
import glob
from pathlib import Path

import numpy as np
import pytest

import helpers as H
from oracle import pipeline as P
from oracle import ref as KR

pytestmark = pytest.mark.gpu
GOLD = Path(__file__).resolve().parent / "golden"
SA = sorted(glob.glob(str(GOLD / "sa_*.npz")))
EXACT_TIES = {"vacancy_fcc", "interstitial_fcc", "slab_fcc", "wire_fcc", "perfect_diamond"}


def _close(ro, ho, check_all_rmsd=True):
    assert np.array_equal(ro[:, 0], ho[:, 0]), "structure types differ"
    assert np.array_equal(ro[:, 1], ho[:, 1]), "alloy ordering differs"
    m = ro[:, 0] > 0
    sel = slice(None) if check_all_rmsd else m
    assert np.allclose(ro[sel, 2], ho[sel, 2], rtol=1e-6, atol=1e-7)
    assert np.allclose(ro[m, 3], ho[m, 3], rtol=1e-6)
    dq = np.minimum(np.abs(ro[m, 4:] - ho[m, 4:]).max(axis=1), np.abs(ro[m, 4:] + ho[m, 4:]).max(axis=1))
    assert dq.size == 0 or dq.max() < 1e-6


@pytest.mark.parametrize("path", SA, ids=[Path(p).stem[3:] for p in SA])
def test_golden_ptm_through_system(path):
    import mdapy_b200 as mp

    d = np.load(path)
    s = mp.System(pos=d["pos"], box=mp.Box(d["box"], boundary=list(d["boundary"])))
    ptm = s.cal_polyhedral_template_matching(return_rmsd=True, return_atomic_distance=True, return_orientation=True)
    assert np.array_equal(np.asarray(s.data["ptm"]), d["ptm"])            # upstream OVITO golden
    _close(d["ref_ptm_output"], ptm.output[: s.N], check_all_rmsd=Path(path).stem[3:] not in EXACT_TIES)
    s2 = mp.System(pos=d["pos"], box=mp.Box(d["box"], boundary=list(d["boundary"])))
    ptm2 = s2.cal_polyhedral_template_matching(structure="all")
    assert np.array_equal(ptm2.output[: s2.N, 0], d["ref_ptm_all_output"][:, 0])     # all eight structures


def _hcp(a=2.95, n=(8, 5, 5)):
    c = a * np.sqrt(8 / 3)
    basis = np.array([[0, 0, 0], [0.5, 0.5, 0], [0.5, 5 / 6, 0.5], [0, 1 / 3, 0.5]])
    cell = np.array([a, a * np.sqrt(3), c])
    ix, iy, iz = np.meshgrid(*[np.arange(k) for k in n], indexing="ij")
    shift = np.stack([ix.ravel(), iy.ravel(), iz.ravel()], 1)
    return ((basis[None] + shift[:, None, :]) * cell).reshape(-1, 3), np.diag(cell * np.array(n))


def _graphene(a=2.46, n=20):
    frac = np.array([[0, 0, 0.5], [0.5, 1 / 6, 0.5], [0.5, 0.5, 0.5], [0, 2 / 3, 0.5]])
    cell = np.array([a, np.sqrt(3) * a, 20.0])
    g = np.stack(np.meshgrid(np.arange(n), np.arange(n), np.arange(1), indexing="ij"), -1).reshape(-1, 3)
    return ((frac[None] + g[:, None]) * cell).reshape(-1, 3), np.diag(cell * np.array([n, n, 1]))


def _seeded():
    rng = np.random.default_rng(5)
    out = []
    p, b = H.fcc(3.615, 9)
    keep = rng.random(p.shape[0]) > 0.02
    out.append(("fcc_vac_rattle", H.rattle(p[keep], 0.08, 1), b, [1, 1, 1], "fcc-hcp-bcc-ico-sc"))
    out.append(("fcc_hot", H.rattle(p, 0.35, 2), b, [1, 1, 1], "default"))
    p2, b2 = H.bcc(2.8665, 11)
    keep = rng.random(p2.shape[0]) > 0.02
    out.append(("bcc_vac_rattle", H.rattle(p2[keep], 0.06, 3), b2, [1, 1, 1], "fcc-hcp-bcc"))
    out.append(("bcc_slab", H.rattle(p2, 0.05, 4), b2, [1, 1, 0], "fcc-hcp-bcc"))
    sc, bs = H.lattice(np.array([[0.0, 0, 0]]), 2.6, 12, 12, 12)
    out.append(("sc_rattle", H.rattle(sc, 0.05, 5), bs, [1, 1, 1], "fcc-hcp-bcc-ico-sc"))
    hp, hb = _hcp()
    out.append(("hcp_rattle", H.rattle(hp, 0.06, 7), hb, [1, 1, 1], "fcc-hcp-bcc-ico-sc"))
    ps, bx = H.shear(H.rattle(p, 0.06, 8), b, xy=0.2, xz=0.1, yz=-0.15)
    out.append(("fcc_triclinic", ps, bx, [1, 1, 1], "fcc-hcp-bcc"))
    g, bg = H.random_gas(2500, 30.0, 6)
    out.append(("gas", g, bg, [1, 1, 1], "all"))
    # two-shell structures: neighbours of neighbours (ptm_multishell.cpp)
    dc, bd = H.diamond(3.567, 6)
    keep = rng.random(dc.shape[0]) > 0.02
    out.append(("dcub_vac_rattle", H.rattle(dc[keep], 0.05, 9), bd, [1, 1, 1], "dcub-dhex"))
    out.append(("dcub_hot_all", H.rattle(dc, 0.3, 10), bd, [1, 1, 1], "all"))
    dh, bh = H.hex_diamond(2.52, 6, 4, 4)
    out.append(("dhex_rattle_slab", H.rattle(dh, 0.06, 11), bh, [1, 1, 0], "all"))
    out.append(("dhex_hot", H.rattle(dh, 0.25, 12), bh, [1, 1, 1], "dcub-dhex-graphene"))
    gr, bgr = _graphene()
    out.append(("graphene_rattle", H.rattle(gr, 0.05, 13), bgr, [1, 1, 0], "graphene"))
    out.append(("graphene_hot_all", H.rattle(gr, 0.2, 14), bgr, [1, 1, 0], "all"))
    out.append(("fcc_hot_all", H.rattle(p, 0.3, 15), b, [1, 1, 1], "all"))
    return out


SEEDED = _seeded()


@pytest.mark.parametrize("case", SEEDED, ids=[c[0] for c in SEEDED])
@pytest.mark.parametrize("thr", [0.1, 10.0])
def test_ptm_matches_reference(case, thr):
    if not KR.available():
        pytest.skip("oracle/_ref not prebuilt")
    from mdapy_b200.device import DeviceSystem

    _, pos, box, bnd, structure = case
    fr = P.Frame(pos, box, bnd)
    f3, idx, _ = P.nearest(KR, fr, 18)
    ro, ri = KR.ptm(structure, *f3.geom(), idx, np.ones(f3.N, np.int32), thr)
    ds = DeviceSystem(0)
    ds.set_atoms(f3.x, f3.y, f3.z, f3.box, f3.origin, f3.boundary)
    ds.put_neighbor(idx, kind=2)
    ho, hi = ds.ptm(structure, thr, np.ones(f3.N, np.int32))
    _close(ro, ho)
    # matched neighbour SETS agree (the order inside follows this library's template point order)
    m = ro[:, 0] > 0
    assert np.array_equal(np.sort(ri[m], axis=1), np.sort(hi[m], axis=1))
    assert np.all(hi[~(ho[:, 2] > 0)] == -1)


def test_ptm_alloy_ordering_b2_l12():
    """Binary orderings: B2 (CsCl) on BCC, L1_2 and L1_0 on FCC, vs the reference."""
    if not KR.available():
        pytest.skip("oracle/_ref not prebuilt")
    from mdapy_b200.device import DeviceSystem

    cases = []
    p, b = H.bcc(2.9, 8)
    cases.append((H.rattle(p, 0.03, 1), b, np.tile([1, 2], p.shape[0] // 2)))                 # B2
    p, b = H.fcc(3.8, 7)
    cases.append((H.rattle(p, 0.03, 2), b, np.tile([2, 1, 1, 1], p.shape[0] // 4)))            # L1_2
    cases.append((H.rattle(p, 0.03, 3), b, np.tile([1, 1, 2, 2], p.shape[0] // 4)))            # L1_0
    cases.append((H.rattle(p, 0.03, 4), b, np.random.default_rng(0).integers(1, 4, p.shape[0])))  # ternary / random
    # zincblende / wurtzite (SiC ordering, 6) and hexagonal BN (7): the two sublattices alternate
    p, b = H.diamond(4.36, 5)
    cases.append((H.rattle(p, 0.03, 5), b, np.tile([1, 1, 1, 1, 2, 2, 2, 2], p.shape[0] // 8)))
    p, b = H.hex_diamond(3.08, 5, 3, 3)
    cases.append((H.rattle(p, 0.03, 6), b, np.tile([1, 2, 2, 1], p.shape[0] // 4)))
    p, b = _graphene(2.5, 16)
    cases.append((H.rattle(p, 0.03, 7), b, np.tile([1, 2, 1, 2], p.shape[0] // 4)))
    seen = set()
    for k, (pos, box, types) in enumerate(cases):
        fr = P.Frame(pos, box, [1, 1, 0] if k == 6 else [1, 1, 1])
        f3, idx, _ = P.nearest(KR, fr, 18)
        t = types.astype(np.int32)
        structure = "fcc-hcp-bcc" if k < 4 else "all"
        ro, _ = KR.ptm(structure, *f3.geom(), idx, t, 0.1)
        ds = DeviceSystem(0)
        ds.set_atoms(f3.x, f3.y, f3.z, f3.box, f3.origin, f3.boundary)
        ds.put_neighbor(idx, kind=2)
        ho, _ = ds.ptm(structure, 0.1, t)
        assert np.array_equal(ro[:, 0], ho[:, 0])
        assert np.array_equal(ro[:, 1], ho[:, 1]), (np.bincount(ro[:, 1].astype(int)), np.bincount(ho[:, 1].astype(int)))
        seen |= set(np.unique(ro[:, 1]).astype(int).tolist())
    assert {1, 2, 3, 4, 5, 6, 7} <= seen | {1}, seen


def test_ptm_host_dropin_and_known_answers():
    import mdapy_b200 as mp
    from mdapy_b200 import _lib as L

    fcc = mp.build_crystal("Cu", "fcc", 3.615, nx=6, ny=6, nz=6)
    fcc.cal_polyhedral_template_matching()
    assert np.all(np.asarray(fcc.data["ptm"]) == 1)
    bcc = mp.build_crystal("Fe", "bcc", 2.8665, nx=7, ny=7, nz=7)
    p = bcc.cal_polyhedral_template_matching(return_rmsd=True)
    assert np.all(np.asarray(bcc.data["ptm"]) == 3) and np.asarray(bcc.data["rmsd"]).max() < 1e-6
    assert np.allclose(p.output[:, 3], 2.8665 * np.sqrt(3) / 2, rtol=1e-9)
    # host-pointer drop-in, called like _ptm.get_ptm
    if KR.available():
        pos, box = H.fcc(3.615, 6)
        pos = H.rattle(pos, 0.07, 11)
        fr = P.Frame(pos, box)
        f3, idx, _ = P.nearest(KR, fr, 18)
        N = f3.N
        b, o, pb = L.box_args(f3.box, f3.origin, f3.boundary)
        t = np.ones(N, np.int32)
        out = np.zeros((N, 8))
        ind = np.zeros((N, 18), np.int32)
        L.check(L.lib().mdb_get_ptm(b"fcc-hcp-bcc", L.dptr(f3.x), L.dptr(f3.y), L.dptr(f3.z), N, L.dptr(b), L.dptr(o),
                                    L.iptr(pb), L.iptr(idx), 18, L.iptr(t), N, 0.1, L.dptr(out), 8, L.iptr(ind), 18, 8))
        ro, _ = KR.ptm("fcc-hcp-bcc", *f3.geom(), idx, t, 0.1)
        _close(ro, out)
    # perfect two-shell crystals (reference tests/test_polyhedral_template_matching.py:34-55)
    dia = mp.System(pos=H.diamond(3.567, 5)[0], box=H.diamond(3.567, 5)[1])
    pd = dia.cal_polyhedral_template_matching(structure="dcub", return_rmsd=True)
    assert np.all(np.asarray(dia.data["ptm"]) == 6) and np.asarray(dia.data["rmsd"]).max() < 1e-6
    assert np.allclose(pd.output[:, 3], 3.567 * np.sqrt(3) / 4, rtol=1e-9)
    lon = mp.System(pos=H.hex_diamond()[0], box=H.hex_diamond()[1])
    lon.cal_polyhedral_template_matching(structure="dcub-dhex")
    assert np.all(np.asarray(lon.data["ptm"]) == 7)
