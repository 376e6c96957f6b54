"""oracle/ref.py -- TEST INFRASTRUCTURE ONLY (never imported by mdapy_b200).

ctypes bindings to ``oracle/_ref/ref_*.so``: mdapy's own C++ translation units
compiled unmodified from ``/root/reference/src`` (see oracle/Makefile).  Each
function mirrors the nanobind entry point it re-exports
with NumPy arrays in / out.  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s cpu_baseline / ``--impl reference`` legs may import this.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_REF_DIR = _HERE / "_ref"
_libs: dict = {}

c_dp = C.POINTER(C.c_double)
c_ip = C.POINTER(C.c_int)


def available() -> bool:
    return (_REF_DIR / "ref_neighbor.so").exists()


def num_threads() -> int:
    return int(os.environ.get("MDAPY_NUM_THREADS", os.cpu_count() or 1))


def _lib(name: str) -> C.CDLL:
    if name not in _libs:
        path = _REF_DIR / f"ref_{name}.so"
        if not path.exists():
            raise FileNotFoundError(
                f"{path} missing: run `make -C oracle ref` where /root/reference exists"
            )
        _libs[name] = C.CDLL(str(path))
    return _libs[name]


def _d(a):
    return a.ctypes.data_as(c_dp)


def _i(a):
    return a.ctypes.data_as(c_ip)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _boxargs(box, origin, boundary):
    b = _f64(box).reshape(3, 3)
    o = _f64(origin).reshape(3)
    p = _i32(boundary).reshape(3)
    return b, o, p


# --------------------------------------------------------------------------
# src/neighbor.cpp
# --------------------------------------------------------------------------
def build_neighbor(x, y, z, box, origin, boundary, rc, max_neigh, nt=None):
    """neighbor.cpp:351 build_neighbor with the neighbor.py:125-129 prefill."""
    x, y, z = _f64(x), _f64(y), _f64(z)
    b, o, p = _boxargs(box, origin, boundary)
    N = x.shape[0]
    verlet = np.full((N, max_neigh), -1, np.int32)
    dist = np.full((N, max_neigh), rc + 1.0, np.float64)
    nn = np.zeros(N, np.int32)
    _lib("neighbor").ref_build_neighbor(
        _d(x), _d(y), _d(z), C.c_int(N), _d(b), _d(o), _i(p), C.c_double(rc),
        _i(verlet), _d(dist), _i(nn), C.c_int(max_neigh), C.c_int(nt or num_threads()))
    return verlet, dist, nn


def build_neighbor_auto(x, y, z, box, origin, boundary, rc, nt=None):
    """neighbor.cpp:189 build_neighbor_without_max_neigh."""
    x, y, z = _f64(x), _f64(y), _f64(z)
    b, o, p = _boxargs(box, origin, boundary)
    N = x.shape[0]
    lib = _lib("neighbor")
    lib.ref_build_neighbor_auto.restype = C.c_int
    pv, pd, pn = c_ip(), c_dp(), c_ip()
    M = lib.ref_build_neighbor_auto(
        _d(x), _d(y), _d(z), C.c_int(N), _d(b), _d(o), _i(p), C.c_double(rc),
        C.byref(pv), C.byref(pd), C.byref(pn), C.c_int(nt or num_threads()))
    verlet = np.ctypeslib.as_array(pv, shape=(N, M)).copy()
    dist = np.ctypeslib.as_array(pd, shape=(N, M)).copy()
    nn = np.ctypeslib.as_array(pn, shape=(N,)).copy()
    lib.ref_free_int(pv)
    lib.ref_free_double(pd)
    lib.ref_free_int(pn)
    return verlet, dist, nn


def sort_verlet_by_distance(verlet, dist, k, nt=None):
    """neighbor.cpp:745 (in place)."""
    assert verlet.flags.c_contiguous and dist.flags.c_contiguous
    N, M = verlet.shape
    _lib("neighbor").ref_sort_verlet_by_distance(
        _i(verlet), _d(dist), C.c_int(N), C.c_int(M), C.c_int(k), C.c_int(nt or num_threads()))


def wrap_positions(x, y, z, box, origin, boundary, nt=None):
    """neighbor.cpp:675 (in place)."""
    b, o, p = _boxargs(box, origin, boundary)
    _lib("neighbor").ref_wrap_positions(
        _d(x), _d(y), _d(z), C.c_int(x.shape[0]), _d(b), _d(o), _i(p), C.c_int(nt or num_threads()))


# --------------------------------------------------------------------------
# src/fast_knn.cpp
# --------------------------------------------------------------------------
def knn(x, y, z, box, origin, boundary, k, nt=None):
    """fast_knn.cpp:846."""
    x, y, z = _f64(x), _f64(y), _f64(z)
    b, o, p = _boxargs(box, origin, boundary)
    N = x.shape[0]
    idx = np.zeros((N, k), np.int32)
    dst = np.zeros((N, k), np.float64)
    _lib("knn").ref_knn(_d(x), _d(y), _d(z), C.c_int(N), _d(b), _d(o), _i(p), C.c_int(k),
                        _i(idx), _d(dst), C.c_int(nt or num_threads()))
    return idx, dst


# --------------------------------------------------------------------------
# src/cna.cpp
# --------------------------------------------------------------------------
def fcna(x, y, z, box, origin, boundary, verlet, nn, rc, nt=None):
    """cna.cpp:429 FixedCNA."""
    x, y, z = _f64(x), _f64(y), _f64(z)
    b, o, p = _boxargs(box, origin, boundary)
    verlet, nn = _i32(verlet), _i32(nn)
    N, M = verlet.shape
    pattern = np.zeros(N, np.int32)
    _lib("cna").ref_fcna(_d(x), _d(y), _d(z), C.c_int(N), _d(b), _d(o), _i(p), _i(verlet), C.c_int(M),
                         _i(nn), _i(pattern), C.c_double(rc), C.c_int(nt or num_threads()))
    return pattern


def acna(x, y, z, box, origin, boundary, verlet, nt=None):
    """cna.cpp:289 AdaptiveCNA."""
    x, y, z = _f64(x), _f64(y), _f64(z)
    b, o, p = _boxargs(box, origin, boundary)
    verlet = _i32(verlet)
    N, M = verlet.shape
    pattern = np.zeros(N, np.int32)
    _lib("cna").ref_acna(_d(x), _d(y), _d(z), C.c_int(N), _d(b), _d(o), _i(p), _i(verlet), C.c_int(M),
                         _i(pattern), C.c_int(nt or num_threads()))
    return pattern


def ids(x, y, z, box, origin, boundary, verlet, nt=None):
    """cna.cpp:163 IdentifyDiamond."""
    x, y, z = _f64(x), _f64(y), _f64(z)
    b, o, p = _boxargs(box, origin, boundary)
    verlet = _i32(verlet)
    N, M = verlet.shape
    new_verlet = np.zeros((N, 12), np.int32)
    pattern = np.zeros(N, np.int32)
    _lib("cna").ref_ids(_d(x), _d(y), _d(z), C.c_int(N), _d(b), _d(o), _i(p), _i(verlet), C.c_int(M),
                        _i(new_verlet), _i(pattern), C.c_int(nt or num_threads()))
    return pattern


# --------------------------------------------------------------------------
# src/centro_symmetry_parameter.cpp, src/ackland_jones_analysis.cpp
# --------------------------------------------------------------------------
def csp(x, y, z, box, origin, boundary, verlet, nnei, nt=None):
    """centro_symmetry_parameter.cpp:12 get_csp."""
    x, y, z = _f64(x), _f64(y), _f64(z)
    b, o, p = _boxargs(box, origin, boundary)
    verlet = _i32(verlet)
    N, M = verlet.shape
    out = np.zeros(N, np.float64)
    _lib("csp").ref_csp(_d(x), _d(y), _d(z), C.c_int(N), _d(b), _d(o), _i(p), _i(verlet), C.c_int(M),
                        C.c_int(nnei), _d(out), C.c_int(nt or num_threads()))
    return out


def aja(x, y, z, box, origin, boundary, verlet, dist, nt=None):
    """ackland_jones_analysis.cpp:9 compute_aja."""
    x, y, z = _f64(x), _f64(y), _f64(z)
    b, o, p = _boxargs(box, origin, boundary)
    verlet, dist = _i32(verlet), _f64(dist)
    N, M = verlet.shape
    out = np.zeros(N, np.int32)
    _lib("aja").ref_aja(_d(x), _d(y), _d(z), C.c_int(N), _d(b), _d(o), _i(p), _i(verlet), C.c_int(M),
                        _d(dist), C.c_int(dist.shape[1]), _i(out), C.c_int(nt or num_threads()))
    return out


# --------------------------------------------------------------------------
# src/polyhedral_template_matching.cpp (+ extern/ptm)
# --------------------------------------------------------------------------
def ptm(structure, x, y, z, box, origin, boundary, verlet, atom_types, rmsd_threshold, nt=None):
    """polyhedral_template_matching.cpp:135 get_ptm -> (output[N,8], ptm_indices[N,18])."""
    x, y, z = _f64(x), _f64(y), _f64(z)
    b, o, p = _boxargs(box, origin, boundary)
    verlet = _i32(verlet)
    types = _i32(atom_types)
    N, M = verlet.shape
    out = np.zeros((N, 8), np.float64)
    ind = np.zeros((N, 18), np.int32)
    _lib("ptm").ref_ptm(structure.encode(), _d(x), _d(y), _d(z), C.c_int(N), _d(b), _d(o), _i(p),
                        _i(verlet), C.c_int(M), _i(types), C.c_int(types.shape[0]),
                        C.c_double(rmsd_threshold), _d(out), C.c_int(8), _i(ind), C.c_int(18),
                        C.c_int(nt or num_threads()))
    return out, ind


# --------------------------------------------------------------------------
# src/steinhardt_bond_orientation.cpp
# --------------------------------------------------------------------------
def get_sq(x, y, z, box, origin, boundary, verlet, dist, nn, llist, nnn=0, rc=-1.0, average=False,
           wl=False, wlhat=False, use_voronoi=False, weight=None, nt=None):
    """steinhardt_bond_orientation.cpp:677 get_sq -> (qnarray, qlm_r, qlm_i).

    ``rc`` follows steinhardt_bond_orientation.py:238-245 (huge rc for nnn / voronoi)."""
    x, y, z = _f64(x), _f64(y), _f64(z)
    b, o, p = _boxargs(box, origin, boundary)
    verlet, dist, nn = _i32(verlet), _f64(dist), _i32(nn)
    ll = _i32(llist)
    N, M = verlet.shape
    lmax = int(ll.max())
    ndeg = ll.shape[0]
    qr = np.zeros((N, ndeg, 2 * lmax + 1))
    qi = np.zeros_like(qr)
    ncol = ndeg * (1 + int(bool(wl)) + int(bool(wlhat)))
    qn = np.zeros((N, ncol))
    if use_voronoi:
        rc = 10000000000.0
    elif nnn > 0:
        rc = 1000000000.0
    use_weight = weight is not None
    w = _f64(weight) if use_weight else np.zeros((2, 2))
    _lib("sbo").ref_get_sq(
        _d(x), _d(y), _d(z), C.c_int(N), _d(b), _d(o), _i(p), _i(verlet), C.c_int(M), _d(dist), _i(nn),
        _d(w), C.c_int(w.shape[0]), C.c_int(w.shape[1]), _i(ll), C.c_int(ndeg), C.c_int(nnn),
        C.c_int(lmax), C.c_int(wl), C.c_int(wlhat), C.c_int(average), C.c_int(use_voronoi),
        C.c_double(rc), C.c_int(use_weight), _d(qr), _d(qi), _d(qn), C.c_int(ncol),
        C.c_int(nt or num_threads()))
    return qn, qr, qi


def solid_liquid(q6index, q6, verlet, dist, nn, qlm_r, qlm_i, threshold, n_bond, nnn=0, rc=-1.0,
                 use_voronoi=False, nt=None):
    """steinhardt_bond_orientation.cpp:578 identifySolidLiquid -> (solidliquid, nbond)."""
    verlet, dist, nn = _i32(verlet), _f64(dist), _i32(nn)
    q6, qlm_r, qlm_i = _f64(q6), _f64(qlm_r), _f64(qlm_i)
    N, M = verlet.shape
    if use_voronoi:
        rc = 10000000000.0
    elif nnn > 0:
        rc = 1000000000.0
    sl = np.zeros(N, np.int32)
    nb = np.zeros(N, np.int32)
    _lib("sbo").ref_solid_liquid(
        C.c_int(q6index), _d(q6), _i(verlet), C.c_int(N), C.c_int(M), _d(dist), _i(nn), _d(qlm_r), _d(qlm_i),
        C.c_int(qlm_r.shape[1]), C.c_int(qlm_r.shape[2]), C.c_double(threshold), C.c_int(n_bond),
        _i(sl), _i(nb), C.c_int(use_voronoi), C.c_int(nnn), C.c_double(rc), C.c_int(nt or num_threads()))
    return sl, nb


# --------------------------------------------------------------------------
# src/radial_distribution_function.cpp
# --------------------------------------------------------------------------
def rdf_list(verlet, dist, nn, type_list, ntype, rc, nbin):
    """radial_distribution_function.cpp:22 _rdf -> counts[T,T,nbin]."""
    verlet, dist, nn, t = _i32(verlet), _f64(dist), _i32(nn), _i32(type_list)
    N, M = verlet.shape
    g = np.zeros((ntype, ntype, nbin))
    _lib("rdf").ref_rdf(_i(verlet), C.c_int(N), C.c_int(M), _d(dist), _i(nn), _i(t), _d(g), C.c_int(ntype),
                        C.c_double(rc), C.c_int(nbin))
    return g


def rdf_single(verlet, dist, nn, rc, nbin):
    """radial_distribution_function.cpp:56 _rdf_single_species -> counts[nbin]."""
    verlet, dist, nn = _i32(verlet), _f64(dist), _i32(nn)
    N, M = verlet.shape
    g = np.zeros(nbin)
    _lib("rdf").ref_rdf_single(_i(verlet), C.c_int(N), C.c_int(M), _d(dist), _i(nn), _d(g), C.c_double(rc),
                               C.c_int(nbin))
    return g


def rdf_streaming(x, y, z, type_list, ntype, box, origin, boundary, rc, nbin, nt=None):
    """radial_distribution_function.cpp:143 _rdf_streaming -> counts[T,T,nbin]."""
    x, y, z = _f64(x), _f64(y), _f64(z)
    b, o, p = _boxargs(box, origin, boundary)
    t = _i32(type_list)
    g = np.zeros((ntype, ntype, nbin))
    _lib("rdf").ref_rdf_streaming(_d(x), _d(y), _d(z), C.c_int(x.shape[0]), _i(t), _d(b), _d(o), _i(p), _d(g),
                                  C.c_int(ntype), C.c_double(rc), C.c_int(nbin), C.c_int(nt or num_threads()))
    return g


# --------------------------------------------------------------------------
# src/repeat_cell.cpp
# --------------------------------------------------------------------------
def repeat_cell(box, pos, nx, ny, nz, nt=None):
    """repeat_cell.cpp:19 -> new_pos[n*nx*ny*nz, 3]."""
    b = _f64(box).reshape(3, 3)
    pos = _f64(pos)
    n = pos.shape[0]
    out = np.zeros(n * nx * ny * nz * 3)
    _lib("repeat").ref_repeat_cell(_d(out), _d(b), _d(pos), C.c_int(n), C.c_int(nx), C.c_int(ny), C.c_int(nz),
                                   C.c_int(nt or num_threads()))
    return out.reshape(-1, 3)


def identify_sftb_fcc(structure_types, ptm_indices12, identify_esf=True, nt=None):
    """identify_fcc_planar_faults.py:71-84 + identify_fcc_planar_faults.cpp:43 -> fault_types int32[N]."""
    st = _i32(structure_types)
    idx = _i32(np.ascontiguousarray(ptm_indices12))
    hcp = np.where(st == 2)[0].astype(np.int32)
    hn = np.zeros((hcp.shape[0], 12), np.int32)
    fault = np.zeros_like(st)
    _lib("fccpft").ref_identify_sftb_fcc(_i(hcp), C.c_int(hcp.shape[0]), _i(hn), _i(idx), _i(st), C.c_int(st.shape[0]),
                                         _i(fault), C.c_int(int(bool(identify_esf))), C.c_int(nt or num_threads()))
    return fault


def chill_plus(x, y, z, box, origin, boundary, verlet, dist, nn, rc, nt=None):
    """chill_plus.cpp:76 compute_chill_plus -> pattern int32[N]."""
    x, y, z = _f64(x), _f64(y), _f64(z)
    b, o, p = _boxargs(box, origin, boundary)
    verlet, dist, nn = _i32(verlet), _f64(dist), _i32(nn)
    pat = np.zeros(x.shape[0], np.int32)
    _lib("chill").ref_compute_chill_plus(_d(x), _d(y), _d(z), C.c_int(x.shape[0]), _d(b), _d(o), _i(p), _i(verlet),
                                         C.c_int(verlet.shape[1]), _d(dist), _i(nn), C.c_double(rc), _i(pat),
                                         C.c_int(nt or num_threads()))
    return pat


def build_bond(verlet, dist, nn, type_list, cutoff_matrix, nt=None):
    """build_bond.cpp:9 -> (Nbond, 2) int32 pairs i < j (row order depends on the OpenMP schedule)."""
    verlet, dist, nn, t = _i32(verlet), _f64(dist), _i32(nn), _i32(type_list)
    cm = _f64(cutoff_matrix)
    out = np.zeros((verlet.size, 2), np.int32)
    lib = _lib("buildbond")
    lib.ref_build_bond.restype = C.c_int
    n = lib.ref_build_bond(_i(verlet), C.c_int(verlet.shape[0]), C.c_int(verlet.shape[1]), _d(dist), _i(nn), _i(t), _d(cm),
                           C.c_int(cm.shape[0]), _i(out), C.c_int(nt or num_threads()))
    return out[:n].copy()


def filter_overlap_atom(x, y, z, box, origin, boundary, rc, nt=None):
    """neighbor.cpp:390 -> keep flags (bool[N]); the higher index of every pair within rc is dropped."""
    x, y, z = _f64(x), _f64(y), _f64(z)
    b, o, p = _boxargs(box, origin, boundary)
    keep = np.zeros(x.shape[0], np.uint8)
    _lib("neighbor").ref_filter_overlap_atom(_d(x), _d(y), _d(z), C.c_int(x.shape[0]), _d(b), _d(o), _i(p),
                                             C.c_double(rc), keep.ctypes.data_as(C.c_void_p), C.c_int(nt or num_threads()))
    return keep.astype(bool)


def transform_and_filter(x, y, z, rotation, center, target, coeffs, nt=None):
    """polycrystal.cpp:21 -> (count, 3) positions of the atoms inside the convex cell, input order."""
    x, y, z = _f64(x), _f64(y), _f64(z)
    R, c, t, pl = _f64(rotation).reshape(3, 3), _f64(center).reshape(3), _f64(target).reshape(3), _f64(coeffs).reshape(-1, 4)
    out = np.zeros((x.shape[0], 3))
    lib = _lib("polycrystal")
    lib.ref_transform_and_filter.restype = C.c_int
    n = lib.ref_transform_and_filter(_d(x), _d(y), _d(z), C.c_int(x.shape[0]), _d(R), _d(c), _d(t), _d(pl),
                                     C.c_int(pl.shape[0]), _d(out), C.c_int(nt or num_threads()))
    return out[:n].copy()


# --------------------------------------------------------------------------
# further list consumers: common neighbour parameter, Warren-Cowley, average_by_neighbor
# --------------------------------------------------------------------------
def cnp(x, y, z, box, origin, boundary, verlet, dist, nn, rc, nt=None):
    """common_neighbor_parameter.cpp:10 compute_cnp."""
    x, y, z = _f64(x), _f64(y), _f64(z)
    b, o, p = _boxargs(box, origin, boundary)
    verlet, dist, nn = _i32(verlet), _f64(dist), _i32(nn)
    N, M = verlet.shape
    out = np.zeros(N, np.float64)
    _lib("cnp").ref_cnp(_d(x), _d(y), _d(z), C.c_int(N), _d(b), _d(o), _i(p), _i(verlet), C.c_int(M), _d(dist),
                        _i(nn), _d(out), C.c_double(rc), C.c_int(nt or num_threads()))
    return out


def wcp(verlet, nn, type_list, ntype, nt=None):
    """warren_cowley_parameter.cpp:9 get_wcp."""
    verlet, nn, type_list = _i32(verlet), _i32(nn), _i32(type_list)
    N, M = verlet.shape
    out = np.zeros((ntype, ntype), np.float64)
    _lib("wcp").ref_wcp(_i(verlet), C.c_int(N), C.c_int(M), _i(nn), _i(type_list), C.c_int(ntype), _d(out),
                        C.c_int(nt or num_threads()))
    return out


def average_by_neighbor(rc, verlet, dist, nn, value, include_self=True, nt=None):
    """neighbor.cpp:704 average_by_neighbor."""
    verlet, dist, nn, value = _i32(verlet), _f64(dist), _i32(nn), _f64(value)
    N, M = verlet.shape
    out = np.zeros(N, np.float64)
    _lib("neighbor").ref_average_by_neighbor(C.c_double(rc), _i(verlet), C.c_int(N), C.c_int(M), _d(dist), _i(nn),
                                             _d(value), _d(out), C.c_int(int(bool(include_self))),
                                             C.c_int(nt or num_threads()))
    return out


def cluster(verlet, nn, dist=None, rc=0.0):
    """cluster.cpp:9 get_cluster (dist given) / :62 get_cluster_by_bond (dist None) -> (ids, count)."""
    verlet, nn = _i32(verlet), _i32(nn)
    N, M = verlet.shape
    out = np.full(N, -1, np.int32)
    if dist is None:
        cnt = _lib("cluster").ref_get_cluster_by_bond(_i(verlet), C.c_int(N), C.c_int(M), _i(nn), _i(out))
    else:
        dist = _f64(dist)
        cnt = _lib("cluster").ref_get_cluster(_i(verlet), C.c_int(N), C.c_int(M), _d(dist), _i(nn), C.c_double(rc), _i(out))
    return out, int(cnt)


def filter_by_type(verlet, dist, nn, type_list, type1, type2, r, nt=None):
    """cluster.cpp:114 filter_by_type -> filtered copy of verlet."""
    v = _i32(verlet).copy()
    dist, nn, type_list = _f64(dist), _i32(nn), _i32(type_list)
    t1, t2, r = _i32(type1), _i32(type2), _f64(r)
    N, M = v.shape
    _lib("cluster").ref_filter_by_type(_i(v), C.c_int(N), C.c_int(M), _d(dist), _i(nn), _i(type_list), _i(t1), _i(t2), _d(r),
           C.c_int(t1.shape[0]), C.c_int(nt or num_threads()))
    return v


def structure_entropy(rc, sigma, use_local_density, volume, dist, nn, nt=None):
    """structure_entropy.cpp:11 calculate_structure_entropy."""
    dist, nn = _f64(dist), _i32(nn)
    N, M = dist.shape
    out = np.zeros(N, np.float64)
    _lib("entropy").ref_structure_entropy(C.c_double(rc), C.c_double(sigma), C.c_int(int(bool(use_local_density))),
                                    C.c_double(volume), _d(dist), C.c_int(N), C.c_int(M), _i(nn), _d(out),
                                    C.c_int(nt or num_threads()))
    return out


def compute_temp(verlet, dist, vx, vy, vz, mass, rc, nt=None):
    """atomic_temperature.cpp:7 compute_temp (velocities A/ps, masses g/mol) -> T [K]."""
    verlet, dist = _i32(verlet), _f64(dist)
    vx, vy, vz, mass = _f64(vx), _f64(vy), _f64(vz), _f64(mass)
    N, M = verlet.shape
    out = np.zeros(N, np.float64)
    _lib("atomtemp").ref_compute_temp(_i(verlet), C.c_int(N), C.c_int(M), _d(dist), _d(vx), _d(vy), _d(vz), _d(mass), _d(out),
                             C.c_double(rc), C.c_int(nt or num_threads()))
    return out


def compute_bond(x, y, z, box, origin, boundary, verlet, dist, nn, rc, nbin, nt=None):
    """bond_analysis.cpp:7 compute_bond -> (bond_length_distribution, bond_angle_distribution)."""
    x, y, z = _f64(x), _f64(y), _f64(z)
    b, o, p = _boxargs(box, origin, boundary)
    verlet, dist, nn = _i32(verlet), _f64(dist), _i32(nn)
    N, M = verlet.shape
    bl, ba = np.zeros(nbin, np.int32), np.zeros(nbin, np.int32)
    _lib("bond").ref_compute_bond(_d(x), _d(y), _d(z), C.c_int(N), _d(b), _d(o), _i(p), _i(verlet), C.c_int(M), _d(dist),
                             _i(nn), _i(bl), _i(ba), C.c_double(rc / nbin), C.c_double(180.0 / nbin), C.c_double(rc),
                             C.c_int(nbin), C.c_int(nt or num_threads()))
    return bl, ba


def compute_adf(x, y, z, box, origin, boundary, verlet, dist, nn, rc_list, pair_list, type_list, nbin, nt=None):
    """bond_analysis.cpp:120 compute_adf -> int32[Npair, nbin]."""
    x, y, z = _f64(x), _f64(y), _f64(z)
    b, o, p = _boxargs(box, origin, boundary)
    verlet, dist, nn = _i32(verlet), _f64(dist), _i32(nn)
    rcl, pl, t = _f64(rc_list), _i32(pair_list), _i32(type_list)
    N, M = verlet.shape
    out = np.zeros((pl.shape[0], nbin), np.int32)
    _lib("bond").ref_compute_adf(_d(x), _d(y), _d(z), C.c_int(N), _d(b), _d(o), _i(p), _i(verlet), C.c_int(M), _d(dist), _i(nn),
                            C.c_double(180.0 / nbin), _d(rcl), _i(pl), C.c_int(pl.shape[0]), _i(t), C.c_int(nbin), _i(out),
                            C.c_int(nt or num_threads()))
    return out


def voronoi_volume(x, y, z, box, origin, boundary, nt=None):
    """voronoi.cpp:16 get_voronoi_volume_number_radius -> (volume f64[N], faces int32[N], cavity_radius f64[N])."""
    x, y, z = _f64(x), _f64(y), _f64(z)
    b, o, p = _boxargs(box, origin, boundary)
    N = x.shape[0]
    vol, nn, rad = np.zeros(N), np.zeros(N, np.int32), np.zeros(N)
    _lib("voronoi").ref_voronoi_volume_number_radius(_d(x), _d(y), _d(z), C.c_int(N), _d(b), _d(o), _i(p), _d(vol),
                                                     _i(nn), _d(rad), C.c_int(nt or num_threads()))
    return vol, nn, rad


def voronoi_neighbor(x, y, z, box, origin, boundary, a_face_area_threshold=-1.0, r_face_area_threshold=-1.0, nt=None):
    """voronoi.cpp:307 get_voronoi_neighbor -> (verlet int32[N,M], dist f64[N,M], face_area f64[N,M], nn int32[N]);
    rows in voro++'s face order, filtered entries are -1 / 10000 / 0 in place."""
    x, y, z = _f64(x), _f64(y), _f64(z)
    b, o, p = _boxargs(box, origin, boundary)
    N = x.shape[0]
    lib = _lib("voronoi")
    lib.ref_voronoi_neighbor.restype = C.c_int
    pv, pd, pa, pn = c_ip(), c_dp(), c_dp(), c_ip()
    M = lib.ref_voronoi_neighbor(_d(x), _d(y), _d(z), C.c_int(N), _d(b), _d(o), _i(p), C.c_double(a_face_area_threshold),
                                 C.c_double(r_face_area_threshold), C.byref(pv), C.byref(pd), C.byref(pa), C.byref(pn),
                                 C.c_int(nt or num_threads()))
    verlet = np.ctypeslib.as_array(pv, shape=(N, M)).copy()
    dist = np.ctypeslib.as_array(pd, shape=(N, M)).copy()
    area = np.ctypeslib.as_array(pa, shape=(N, M)).copy()
    nn = np.ctypeslib.as_array(pn, shape=(N,)).copy()
    lib.ref_voro_free_int(pv)
    lib.ref_voro_free_double(pd)
    lib.ref_voro_free_double(pa)
    lib.ref_voro_free_int(pn)
    return verlet, dist, area, nn


def _voronoi_tri_setup(box, boundary):
    """src/mdapy/voronoi.py:140-152 + box.py:425-443: rotate the cell into LAMMPS form, triple the open axes."""
    box = np.asarray(box, float)[:3]
    need_rotation = bool(abs(box[0, 1]) > 1e-10 or abs(box[0, 2]) > 1e-10 or abs(box[1, 2]) > 1e-10
                         or box[0, 0] < 0 or box[1, 1] < 0 or box[2, 2] < 0)
    ax = np.linalg.norm(box[0])
    bx = box[1] @ (box[0] / ax)
    by = np.sqrt(np.linalg.norm(box[1]) ** 2 - bx ** 2)
    cx = box[2] @ (box[0] / ax)
    cy = (box[1] @ box[2] - bx * cx) / by
    cz = np.sqrt(np.linalg.norm(box[2]) ** 2 - cx ** 2 - cy ** 2)
    aligned = np.array([[ax, bx, cx], [0, by, cy], [0, 0, cz]], dtype=np.float64).T
    rotation = np.linalg.solve(box, aligned)
    for i in range(3):
        if boundary[i] == 0:
            aligned[i] *= 3
    return np.ascontiguousarray(aligned), np.ascontiguousarray(rotation), need_rotation


def voronoi_volume_tri(x, y, z, box, origin, boundary, nt=None):
    """voronoi.py:303-322 -> voronoi.cpp:73 get_voronoi_volume_number_radius_tri (triclinic cells)."""
    x, y, z = _f64(x), _f64(y), _f64(z)
    aligned, rotation, need = _voronoi_tri_setup(box, boundary)
    b, o, p = _boxargs(aligned, origin, boundary)
    N = x.shape[0]
    vol, nn, rad = np.zeros(N), np.zeros(N, np.int32), np.zeros(N)
    _lib("voronoi").ref_voronoi_volume_number_radius_tri(_d(x), _d(y), _d(z), C.c_int(N), _d(b), _d(o), _i(p),
                                                         _d(rotation), _d(vol), _i(nn), _d(rad), C.c_int(int(need)),
                                                         C.c_int(nt or num_threads()))
    return vol, nn, rad


def voronoi_neighbor_tri(x, y, z, box, origin, boundary, a_face_area_threshold=-1.0, r_face_area_threshold=-1.0, nt=None):
    """voronoi.py:140-166 -> voronoi.cpp:149 get_voronoi_neighbor_tri.  NOTE the reference takes minimum images of the
    UNROTATED coordinate differences in the ROTATED cell: distances are only meaningful when no rotation is needed."""
    x, y, z = _f64(x), _f64(y), _f64(z)
    aligned, rotation, need = _voronoi_tri_setup(box, boundary)
    b, o, p = _boxargs(aligned, origin, boundary)
    N = x.shape[0]
    lib = _lib("voronoi")
    lib.ref_voronoi_neighbor_tri.restype = C.c_int
    pv, pd, pa, pn = c_ip(), c_dp(), c_dp(), c_ip()
    M = lib.ref_voronoi_neighbor_tri(_d(x), _d(y), _d(z), C.c_int(N), _d(b), _d(o), _i(p), _d(rotation),
                                     C.c_int(int(need)), C.c_double(a_face_area_threshold),
                                     C.c_double(r_face_area_threshold), C.byref(pv), C.byref(pd), C.byref(pa),
                                     C.byref(pn), C.c_int(nt or num_threads()))
    verlet = np.ctypeslib.as_array(pv, shape=(N, M)).copy()
    dist = np.ctypeslib.as_array(pd, shape=(N, M)).copy()
    area = np.ctypeslib.as_array(pa, shape=(N, M)).copy()
    nn = np.ctypeslib.as_array(pn, shape=(N,)).copy()
    lib.ref_voro_free_int(pv)
    lib.ref_voro_free_double(pd)
    lib.ref_voro_free_double(pa)
    lib.ref_voro_free_int(pn)
    return verlet, dist, area, nn
