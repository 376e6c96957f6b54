"""GPU parity: Steinhardt q_l / w_l / averaging / solid-liquid and the RDF kernels against the oracle,
through the device handle AND through the host-pointer C-ABI drop-ins (section A of the header).
Floating-point bar from north_star: 1e-6 relative; what is asserted here is stronger (bit equality of
q_lm / q_l for a given list, exact integer pair counts)."""
import ctypes as C

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
    out.append(("fcc_rattled", H.rattle(p, 0.06, 0), b, [1, 1, 1], 3.3))
    out.append(("fcc_hot_slab", H.rattle(p, 0.3, 1), b, [1, 1, 0], 4.1))
    ps, bs = H.shear(H.rattle(p, 0.05, 2), b, xy=0.25, xz=-0.1, yz=0.3)
    out.append(("triclinic", ps, bs, [1, 1, 1], 3.4))
    g, bg = H.random_gas(2500, 28.0, 3)
    out.append(("gas_open", g, bg, [0, 1, 0], 3.5))
    return out


CASES = _cases()


def _bits(a):
    return np.nan_to_num(np.ascontiguousarray(a), nan=-7.0).view(np.int64)


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_steinhardt_and_solid_liquid(case):
    _, pos, box, bnd, rc = case
    x, y, z = (np.ascontiguousarray(pos[:, k]) for k in range(3))
    o = np.zeros(3)
    rv, rd, rn = K.build_neighbor_auto(x, y, z, box, o, bnd, rc)
    ds = _dev()
    ds.set_atoms(x, y, z, box, o, bnd)
    ds.build_neighbor(rc)
    for kw in (dict(), dict(average=True), dict(wl=True, wlhat=True), dict(average=True, wl=True)):
        rq = K.get_sq(x, y, z, box, o, bnd, rv, rd, rn, [4, 6], rc=rc, **kw)
        gq = ds.steinhardt([4, 6], rc=rc, fetch_qlm=True, **kw)
        for a, b in zip(rq, gq):
            assert np.array_equal(_bits(a), _bits(b)), kw
    # larger degrees go through the global-memory accumulator path
    rq = K.get_sq(x, y, z, box, o, bnd, rv, rd, rn, [4, 6, 8, 10, 12], rc=rc, wl=True)
    gq = ds.steinhardt([4, 6, 8, 10, 12], rc=rc, wl=True, fetch_qlm=True)
    assert np.allclose(np.nan_to_num(rq[0]), np.nan_to_num(gq[0]), rtol=1e-12, atol=1e-14)
    assert np.array_equal(_bits(rq[1]), _bits(gq[1]))
    # solid / liquid on q6
    rq = K.get_sq(x, y, z, box, o, bnd, rv, rd, rn, [4, 6], rc=rc)
    ds.steinhardt([4, 6], rc=rc)
    rs = K.solid_liquid(1, np.ascontiguousarray(rq[0][:, 1]), rv, rd, rn, rq[1], rq[2], 0.7, 7, rc=rc)
    gs = ds.solid_liquid(1, 0.7, 7, rc=rc)
    assert np.array_equal(rs[0], gs[0]) and np.array_equal(rs[1], gs[1])
    # nnn source on a kNN list
    ri, rdd = K.knn(x, y, z, box, o, bnd, 12)
    rq = K.get_sq(x, y, z, box, o, bnd, ri, rdd, np.full(x.shape[0], 12, np.int32), [6], nnn=12)
    ds.put_neighbor(ri, rdd, np.full(x.shape[0], 12, np.int32), kind=2)
    gq = ds.steinhardt([6], nnn=12)
    assert np.array_equal(_bits(rq[0]), _bits(gq[0]))


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_rdf_counts_exact(case):
    _, pos, box, bnd, rc = case
    x, y, z = (np.ascontiguousarray(pos[:, k]) for k in range(3))
    o = np.zeros(3)
    rv, rd, rn = K.build_neighbor_auto(x, y, z, box, o, bnd, rc)
    t = (np.arange(x.shape[0]) % 3).astype(np.int32)
    ds = _dev()
    ds.set_atoms(x, y, z, box, o, bnd)
    ds.build_neighbor(rc)
    for nbin in (40, 500, 5000):
        assert np.array_equal(ds.rdf_counts(rc, nbin, t, 3), K.rdf_list(rv, rd, rn, t, 3, rc, nbin))
        assert np.array_equal(ds.rdf_counts(rc, nbin, None, 1), K.rdf_single(rv, rd, rn, rc, nbin))
        assert np.array_equal(ds.rdf_counts(rc, nbin, t, 3, streaming=True),
                              K.rdf_streaming(x, y, z, t, 3, box, o, bnd, rc, nbin))
    # smaller cut-off than the cached list: list kernel filters by distance
    assert np.array_equal(ds.rdf_counts(rc * 0.8, 64, t, 3), K.rdf_list(rv, rd, rn, t, 3, rc * 0.8, 64))


def test_host_pointer_dropins():
    """Section A entry points called exactly like the nanobind functions they replace."""
    from mdapy_b200 import _lib as L

    lib = L.lib()
    _, pos, box, bnd, rc = CASES[0]
    x, y, z = (np.ascontiguousarray(pos[:, k]) for k in range(3))
    N = x.shape[0]
    b, o, p = L.box_args(box, np.zeros(3), bnd)
    rv, rd, rn = K.build_neighbor(x, y, z, box, o, bnd, rc, 40)
    v = np.full((N, 40), -1, np.int32)
    d = np.full((N, 40), rc + 1.0)
    n = np.zeros(N, np.int32)
    L.check(lib.mdb_build_neighbor(L.dptr(x), L.dptr(y), L.dptr(z), N, L.dptr(b), L.dptr(o), L.iptr(p), rc,
                                   L.iptr(v), L.dptr(d), L.iptr(n), 40, 8))
    assert np.array_equal(v, rv) and np.array_equal(d, rd) and np.array_equal(n, rn)
    h, M = C.c_void_p(), C.c_int()
    L.check(lib.mdb_build_neighbor_without_max_neigh(L.dptr(x), L.dptr(y), L.dptr(z), N, L.dptr(b), L.dptr(o),
                                                     L.iptr(p), rc, 8, C.byref(h), C.byref(M)))
    av, ad, an = K.build_neighbor_auto(x, y, z, box, o, bnd, rc)
    v2 = np.empty((N, M.value), np.int32)
    d2 = np.empty((N, M.value))
    n2 = np.empty(N, np.int32)
    L.check(lib.mdb_neighbor_auto_fetch(h, L.iptr(v2), L.dptr(d2), L.iptr(n2)))
    assert np.array_equal(v2, av) and np.array_equal(d2, ad) and np.array_equal(n2, an)
    # sort
    K.sort_verlet_by_distance(av, ad, 12)
    L.check(lib.mdb_sort_verlet_by_distance(L.iptr(v2), L.dptr(d2), N, M.value, 12, 8))
    assert np.array_equal(v2, av) and np.array_equal(d2, ad)
    # fcna / acna / csp / aja / knn
    pat = np.zeros(N, np.int32)
    L.check(lib.mdb_fcna(L.dptr(x), L.dptr(y), L.dptr(z), N, L.dptr(b), L.dptr(o), L.iptr(p), L.iptr(rv), 40,
                         L.iptr(rn), L.iptr(pat), rc, 8))
    assert np.array_equal(pat, K.fcna(x, y, z, box, o, bnd, rv, rn, rc))
    ki, kd = K.knn(x, y, z, box, o, bnd, 14)
    gi, gd = np.zeros((N, 14), np.int32), np.zeros((N, 14))
    L.check(lib.mdb_knn(L.dptr(x), L.dptr(y), L.dptr(z), N, L.dptr(b), L.dptr(o), L.iptr(p), 14, L.iptr(gi),
                        L.dptr(gd), 8))
    assert np.array_equal(gd, kd)
    L.check(lib.mdb_acna(L.dptr(x), L.dptr(y), L.dptr(z), N, L.dptr(b), L.dptr(o), L.iptr(p), L.iptr(ki), 14,
                         L.iptr(pat), 8))
    assert np.array_equal(pat, K.acna(x, y, z, box, o, bnd, ki))
    csp = np.zeros(N)
    L.check(lib.mdb_get_csp(L.dptr(x), L.dptr(y), L.dptr(z), N, L.dptr(b), L.dptr(o), L.iptr(p), L.iptr(ki), 14, 12,
                            L.dptr(csp), 8))
    assert np.array_equal(csp, K.csp(x, y, z, box, o, bnd, ki, 12))
    aja = np.zeros(N, np.int32)
    L.check(lib.mdb_compute_aja(L.dptr(x), L.dptr(y), L.dptr(z), N, L.dptr(b), L.dptr(o), L.iptr(p), L.iptr(ki), 14,
                                L.dptr(kd), 14, L.iptr(aja), 8))
    assert np.array_equal(aja, K.aja(x, y, z, box, o, bnd, ki, kd))
    # get_sq + identifySolidLiquid
    ll = np.array([4, 6], np.int32)
    qr = np.zeros((N, 2, 13))
    qi = np.zeros_like(qr)
    qn = np.zeros((N, 2))
    w = np.zeros((2, 2))
    L.check(lib.mdb_get_sq(L.dptr(x), L.dptr(y), L.dptr(z), N, L.dptr(b), L.dptr(o), L.iptr(p), L.iptr(rv), 40,
                           L.dptr(rd), L.iptr(rn), L.dptr(w), L.iptr(ll), 2, 0, 6, 0, 0, 0, 0, rc, 0, L.dptr(qr),
                           L.dptr(qi), L.dptr(qn), 2, 8))
    rq = K.get_sq(x, y, z, box, o, bnd, rv, rd, rn, [4, 6], rc=rc)
    assert np.array_equal(_bits(qn), _bits(rq[0])) and np.array_equal(_bits(qr), _bits(rq[1]))
    sl, nb = np.zeros(N, np.int32), np.zeros(N, np.int32)
    q6 = np.ascontiguousarray(qn[:, 1])
    L.check(lib.mdb_identify_solid_liquid(1, L.dptr(q6), L.iptr(rv), N, 40, L.dptr(rd), L.iptr(rn), L.dptr(qr),
                                          L.dptr(qi), 2, 13, 0.7, 7, L.iptr(sl), L.iptr(nb), 0, 0, rc, 8))
    rs = K.solid_liquid(1, q6, rv, rd, rn, qr, qi, 0.7, 7, rc=rc)
    assert np.array_equal(sl, rs[0]) and np.array_equal(nb, rs[1])
    # rdf trio
    t = (np.arange(N) % 2).astype(np.int32)
    g = np.zeros((2, 2, 50))
    L.check(lib.mdb_rdf(L.iptr(rv), N, 40, L.dptr(rd), L.iptr(rn), L.iptr(t), L.dptr(g), 2, rc, 50))
    assert np.array_equal(g, K.rdf_list(rv, rd, rn, t, 2, rc, 50))
    g1 = np.zeros(50)
    L.check(lib.mdb_rdf_single_species(L.iptr(rv), N, 40, L.dptr(rd), L.iptr(rn), L.dptr(g1), rc, 50))
    assert np.array_equal(g1, K.rdf_single(rv, rd, rn, rc, 50))
    g2 = np.zeros((2, 2, 50))
    L.check(lib.mdb_rdf_streaming(L.dptr(x), L.dptr(y), L.dptr(z), N, L.iptr(t), L.dptr(b), L.dptr(o), L.iptr(p),
                                  L.dptr(g2), 2, rc, 50, 8))
    assert np.array_equal(g2, K.rdf_streaming(x, y, z, t, 2, box, o, bnd, rc, 50))


def test_error_mapping():
    from mdapy_b200 import _lib as L

    ds = _dev()
    with pytest.raises(RuntimeError):
        ds.fcna(3.0)                         # no list yet
    p, b = H.fcc(3.615, 4)
    with pytest.raises(RuntimeError, match="volume of the box is zero"):
        ds.set_atoms(p[:, 0], p[:, 1], p[:, 2], np.array([[1.0, 1, 0], [2, 2, 0], [0, 0, 1]]), np.zeros(3), [1, 1, 1])
    ds.set_atoms(p[:, 0], p[:, 1], p[:, 2], b, np.zeros(3), [1, 1, 1])
    with pytest.raises(ValueError):
        ds.build_neighbor(-1.0)
    with pytest.raises(ValueError):
        ds.build_knn(25)


def test_small_integer_division():
    """The Legendre recurrences divide by i - m in 1..24; the device uses a reciprocal + two FMA residual steps
    instead of the IEEE division sequence.  It must return the correctly rounded quotient, always: compare
    against a / d on 2^22 values per divisor spanning magnitudes, signs, exact multiples and special values."""
    import ctypes as C

    from mdapy_b200 import _lib as L
    from mdapy_b200.device import DeviceSystem

    ds = DeviceSystem(0)
    rng = np.random.default_rng(7)
    n = 1 << 22
    base = np.concatenate([
        rng.standard_normal(n // 4) * 10.0 ** rng.integers(-12, 12, n // 4),
        rng.random(n // 4) * 3.0,                                   # the actual range of P_l^m intermediates
        np.ldexp(rng.random(n // 4) + 0.5, rng.integers(-1000, 1000, n // 4)),       # normal range only
        rng.integers(-10 ** 9, 10 ** 9, n // 4 - 8).astype(np.float64),
        np.array([0.0, -0.0, np.inf, -np.inf, np.nan, 1e-300, -1.7976931348623157e308, 2.2250738585072014e-308 * 64]),
    ])
    for d in range(1, 26):
        with np.errstate(over="ignore"):
            a = np.ascontiguousarray(base * (d if d % 3 == 0 else 1.0))   # every third divisor: many exact quotients
        bad = C.c_longlong(-1)
        L.check(L.lib().mdb_system_check_small_division(ds._h, L.dptr(a), a.shape[0], d, C.byref(bad)))
        assert bad.value == 0, f"d={d}: {bad.value} quotients differ from IEEE division"
