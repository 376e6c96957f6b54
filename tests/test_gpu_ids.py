"""GPU parity: diamond structure identification (src/cna.cpp:163-287) against the oracle and the
upstream golden labels.  Integer labels: exact equality."""
import ctypes as C
import glob
import os

import numpy as np
import pytest

import helpers as H
from oracle import checker as K
from oracle import pipeline as P

pytestmark = pytest.mark.gpu

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "sa_*.npz")))


def _mixed_blocks():
    """Cubic and hexagonal diamond blocks side by side in an open box: interface atoms are first /
    second neighbours of both kinds, which exercises the lowest-index-wins rule of the sweeps."""
    pc, bc = H.diamond(3.567, 4)
    ph, bh = H.hex_diamond(2.522, 5, 3, 3)
    ph = ph + np.array([bc[0, 0] + 0.9, 0.3, 0.2])
    pos = np.concatenate([pc, ph])
    rng = np.random.default_rng(5)
    pos = pos[rng.permutation(len(pos))]          # interleave the indices of the two phases
    pos = H.rattle(pos, 0.03, 6)
    L = pos.max(0) - pos.min(0) + 4.0
    return pos - pos.min(0) + 2.0, np.diag(L)


def _cases():
    out = []
    p, b = H.diamond(3.567, 5)
    out.append(("cubic_perfect", p, b, [1, 1, 1]))
    rng = np.random.default_rng(0)
    keep = rng.random(len(p)) > 0.03
    out.append(("cubic_vacancies_rattled", H.rattle(p[keep], 0.05, 1), b, [1, 1, 1]))
    q = p.copy()
    m = q[:, 0] < b[0, 0] * 0.3
    q[m] = rng.random((int(m.sum()), 3)) * [b[0, 0] * 0.3, b[1, 1], b[2, 2]]
    out.append(("cubic_plus_disordered_slab", q, b, [1, 1, 0]))
    ph, bh = H.hex_diamond()
    out.append(("hex_perfect", ph, bh, [1, 1, 1]))
    keep = rng.random(len(ph)) > 0.04
    out.append(("hex_vacancies_rattled", H.rattle(ph[keep], 0.04, 2), bh, [1, 1, 1]))
    pm, bm = _mixed_blocks()
    out.append(("mixed_blocks_open", pm, bm, [0, 0, 0]))
    ps, bs = H.shear(H.rattle(p, 0.04, 3), b, xy=0.2, xz=-0.1, yz=0.15)
    out.append(("cubic_triclinic", ps, bs, [1, 1, 1]))
    small, bsm = H.diamond(3.567, 2)              # thickness < 15: replicated internally
    out.append(("cubic_small_box", small, bsm, [1, 1, 1]))
    return out


CASES = _cases()


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_system_matches_oracle(case):
    import mdapy_b200 as mp

    _, pos, box, bnd = case
    want = P.cal_ids(K, P.Frame(pos, box, bnd))
    s = mp.System(pos=pos, box=mp.Box(box, boundary=bnd))
    s.cal_identify_diamond_structure()
    got = np.asarray(s.data["ids"])
    assert got.dtype == np.int32 and np.array_equal(got, want), (np.bincount(got, minlength=7), np.bincount(want, minlength=7))


def test_label_mix_is_nontrivial():
    hist = np.zeros(7, int)
    for _, pos, box, bnd in CASES:
        hist += np.bincount(P.cal_ids(K, P.Frame(pos, box, bnd)), minlength=7)
    assert (hist > 0).all(), hist


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[3:-4] for p in GOLDEN])
def test_upstream_golden(path):
    import mdapy_b200 as mp

    d = np.load(path)
    s = mp.System(pos=d["pos"], box=mp.Box(d["box"], boundary=d["boundary"]))
    s.cal_identify_diamond_structure()
    assert np.array_equal(np.asarray(s.data["ids"]), d["ids"])


def test_cached_cutoff_list_is_reused():
    import mdapy_b200 as mp

    _, pos, box, bnd = CASES[1]
    want = P.cal_ids(K, P.Frame(pos, box, bnd))
    s = mp.System(pos=pos, box=mp.Box(box, boundary=bnd))
    s.build_neighbor(3.2)                         # every atom keeps >= 4 neighbours inside 3.2
    assert s.neighbor_number.min() >= 4
    s.cal_identify_diamond_structure()
    assert np.array_equal(np.asarray(s.data["ids"]), want)


def test_host_pointer_dropin():
    from mdapy_b200 import _lib as L

    lib = L.lib()
    _, pos, box, bnd = CASES[5]
    x, y, z = (np.ascontiguousarray(pos[:, k]) for k in range(3))
    o = np.zeros(3)
    b, oo, p = L.box_args(box, o, bnd)
    idx, _ = K.knn(x, y, z, box, o, bnd, 4)
    want = K.ids(x, y, z, box, o, bnd, idx)
    N = len(x)
    pat = np.full(N, -5, np.int32)
    second = np.full((N, 12), -5, np.int32)
    v = np.ascontiguousarray(idx, np.int32)
    L.check(lib.mdb_ids(L.dptr(x), L.dptr(y), L.dptr(z), N, L.dptr(b), L.dptr(oo), L.iptr(p), L.iptr(v), 4,
                        L.iptr(second), L.iptr(pat), 1))
    assert np.array_equal(pat, want)
    # second-shell list = first three non-self entries of each first neighbour's row
    i = 17
    exp = [k for j in idx[i] for k in [t for t in idx[j] if t != i][:3]]
    assert second[i].tolist() == exp
    L.check(lib.mdb_ids(L.dptr(x), L.dptr(y), L.dptr(z), N, L.dptr(b), L.dptr(oo), L.iptr(p), L.iptr(v), 4,
                        None, L.iptr(pat), 1))
    assert np.array_equal(pat, want)
    with pytest.raises(ValueError):
        L.check(lib.mdb_ids(L.dptr(x), L.dptr(y), L.dptr(z), N, L.dptr(b), L.dptr(oo), L.iptr(p), L.iptr(v[:, :3].copy()),
                            3, None, L.iptr(pat), 1))
