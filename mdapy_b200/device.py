"""Device-resident system handle (section B of include/mdapy_b200.h).

`DeviceSystem` owns one ``mdb_system``: coordinates, the cell-sorted copy and
the neighbour list live in HBM between calls, so chained ``cal_*`` calls do
not round-trip over PCIe.  NumPy views are materialised only when asked for.
"""
from __future__ import annotations

import ctypes as C
import threading

import numpy as np

from . import _lib as L

LIST_NONE, LIST_CUTOFF, LIST_KNN = 0, 1, 2


class DeviceSystem:
    def __init__(self, device: int = 0):
        self._lib = L.lib()
        h = L.c_vp()
        L.check(self._lib.mdb_system_create(int(device), C.byref(h)))
        self._h = h
        self.device = int(device)
        self.N = 0        # atoms on the device (owned + ghosts)
        self.n_rows = 0   # rows of lists / per-atom outputs (owned atoms)
        self.M = 0
        self.max_count = 0
        self._keep = None  # keeps borrowed device tensors / host arrays alive

    def close(self):
        if getattr(self, "_h", None):
            self._lib.mdb_system_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- atoms ---------------------------------------------------------------
    def set_atoms(self, x, y, z, box, origin, boundary):
        x, y, z = L.f64(x), L.f64(y), L.f64(z)
        b, o, p = L.box_args(box, origin, boundary)
        L.check(self._lib.mdb_system_set_atoms(self._h, L.dptr(x), L.dptr(y), L.dptr(z), x.shape[0],
                                               L.dptr(b), L.dptr(o), L.iptr(p)))
        self.N = self.n_rows = x.shape[0]
        self._keep = (x, y, z)
        self.M = 0

    def set_atoms_device(self, dx, dy, dz, box, origin, boundary, stream=None):
        """Borrow torch CUDA tensors (f64, contiguous) without copying."""
        b, o, p = L.box_args(box, origin, boundary)
        if stream is not None:
            L.check(self._lib.mdb_system_set_stream(self._h, C.c_void_p(int(stream))))
        n = int(dx.shape[0])
        L.check(self._lib.mdb_system_set_atoms_device(
            self._h, C.c_void_p(dx.data_ptr()), C.c_void_p(dy.data_ptr()), C.c_void_p(dz.data_ptr()), n,
            L.dptr(b), L.dptr(o), L.iptr(p)))
        self.N = self.n_rows = n
        self._keep = (dx, dy, dz)
        self.M = 0

    def set_slab_device(self, dx, dy, dz, dgid, n_owned, plane0, nplanes, box, origin, boundary, stream=None):
        """Decomposed frame: torch CUDA tensors of owned atoms followed by ghosts, int32 global ids,
        and the stored window of global x cell planes (see mdb_system_set_slab_device)."""
        b, o, p = L.box_args(box, origin, boundary)
        if stream is not None:
            L.check(self._lib.mdb_system_set_stream(self._h, C.c_void_p(int(stream))))
        n = int(dx.shape[0])
        L.check(self._lib.mdb_system_set_slab_device(
            self._h, C.c_void_p(dx.data_ptr()), C.c_void_p(dy.data_ptr()), C.c_void_p(dz.data_ptr()),
            C.c_void_p(dgid.data_ptr()), n, int(n_owned), int(plane0), int(nplanes), L.dptr(b), L.dptr(o), L.iptr(p)))
        self.N, self.n_rows = n, int(n_owned)
        self._keep = (dx, dy, dz, dgid)
        self.M = 0

    def set_local_fraction(self, fraction: float):
        """Decomposed frame: fraction of the box volume the local atoms occupy (density hint only)."""
        L.check(self._lib.mdb_system_set_local_fraction(self._h, float(fraction)))

    def neighbor_device(self):
        """Raw device pointers (int) of verlet / distance / count arrays and the row width."""
        v, d, n, M = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_int(0)
        L.check(self._lib.mdb_system_neighbor_device(self._h, C.byref(v), C.byref(d), C.byref(n), C.byref(M)))
        return v.value, d.value, n.value, M.value

    def synchronize(self):
        L.check(self._lib.mdb_system_synchronize(self._h))

    # -- neighbour list ------------------------------------------------------
    def build_neighbor(self, rc: float, max_neigh=None):
        M, mx = C.c_int(0), C.c_int(0)
        L.check(self._lib.mdb_system_build_neighbor(self._h, float(rc), int(max_neigh or 0), C.byref(M), C.byref(mx)))
        self.M, self.max_count = M.value, mx.value
        return self.M, self.max_count

    def build_knn(self, k: int):
        L.check(self._lib.mdb_system_build_knn(self._h, int(k)))
        self.M = self.max_count = int(k)

    def sort_neighbor(self, k: int):
        L.check(self._lib.mdb_system_sort_neighbor(self._h, int(k)))

    def min_count(self) -> int:
        v = C.c_int(0)
        L.check(self._lib.mdb_system_neighbor_min_count(self._h, C.byref(v)))
        return v.value

    def fetch_neighbor(self, want_verlet=True, want_dist=True, want_nn=True):
        verlet = L.result_empty((self.n_rows, self.M), np.int32) if want_verlet else None
        dist = L.result_empty((self.n_rows, self.M), np.float64) if want_dist else None
        nn = L.result_empty(self.n_rows, np.int32) if want_nn else None
        L.check(self._lib.mdb_system_fetch_neighbor(
            self._h, L.iptr(verlet) if want_verlet else None, L.dptr(dist) if want_dist else None,
            L.iptr(nn) if want_nn else None))
        return verlet, dist, nn

    def put_neighbor(self, verlet, dist=None, nn=None, rc=-1.0, kind=LIST_CUTOFF):
        verlet = L.i32(verlet)
        assert verlet.ndim == 2 and verlet.shape[0] == self.n_rows
        d = L.f64(dist) if dist is not None else None
        n = L.i32(nn) if nn is not None else None
        L.check(self._lib.mdb_system_put_neighbor(
            self._h, L.iptr(verlet), L.dptr(d) if d is not None else None, L.iptr(n) if n is not None else None,
            verlet.shape[1], float(rc), int(kind)))
        self.M = verlet.shape[1]

    # -- descriptors ---------------------------------------------------------
    def fcna(self, rc: float, fetch=True):
        out = L.result_empty(self.n_rows, np.int32) if fetch else None
        L.check(self._lib.mdb_system_fcna(self._h, float(rc), L.iptr(out) if fetch else None))
        return out

    def fused_cna(self, rc: float, fetch=True):
        """Neighbour search + fixed-cutoff CNA without a list in HBM.  Returns (labels or None, used):
        ``used`` is False when the frame is not eligible and nothing was computed."""
        out = L.result_empty(self.n_rows, np.int32) if fetch else None
        used = C.c_int(0)
        L.check(self._lib.mdb_system_fused_cna(self._h, float(rc), L.iptr(out) if fetch else None, C.byref(used)))
        return (out if used.value else None), bool(used.value)

    def acna(self, fetch=True):
        out = L.result_empty(self.n_rows, np.int32) if fetch else None
        L.check(self._lib.mdb_system_acna(self._h, L.iptr(out) if fetch else None))
        return out

    def ids(self, fetch=True):
        out = L.result_empty(self.n_rows, np.int32) if fetch else None
        L.check(self._lib.mdb_system_ids(self._h, L.iptr(out) if fetch else None))
        return out

    def csp(self, nnei: int, fetch=True):
        out = L.result_empty(self.n_rows, np.float64) if fetch else None
        L.check(self._lib.mdb_system_csp(self._h, int(nnei), L.dptr(out) if fetch else None))
        return out

    def aja(self, fetch=True):
        out = L.result_empty(self.n_rows, np.int32) if fetch else None
        L.check(self._lib.mdb_system_aja(self._h, L.iptr(out) if fetch else None))
        return out

    def steinhardt(self, llist, nnn=0, rc=-1.0, average=False, wl=False, wlhat=False, use_voronoi=False,
                   weight=None, fetch_qlm=False, fetch=True):
        ll = L.i32(llist)
        ndeg = ll.shape[0]
        lmax = int(ll.max())
        ncol = ndeg * (1 + int(bool(wl)) + int(bool(wlhat)))
        qn = L.result_empty((self.n_rows, ncol), np.float64) if fetch else None
        qr = L.result_empty((self.n_rows, ndeg, 2 * lmax + 1), np.float64) if fetch_qlm else None
        qi = L.result_empty(qr.shape, np.float64) if fetch_qlm else None
        w = L.f64(weight) if weight is not None else None
        if w is not None:
            assert w.shape == (self.n_rows, self.M)
        L.check(self._lib.mdb_system_steinhardt(
            self._h, L.iptr(ll), ndeg, int(nnn), float(rc), int(bool(average)), int(bool(wl)), int(bool(wlhat)),
            int(bool(use_voronoi)), L.dptr(w) if w is not None else None, L.dptr(qn) if fetch else None,
            L.dptr(qr) if fetch_qlm else None, L.dptr(qi) if fetch_qlm else None))
        return qn, qr, qi

    def solid_liquid(self, q6index, threshold, n_bond, nnn=0, rc=-1.0, use_voronoi=False):
        sl = L.result_empty(self.n_rows, np.int32)
        nb = L.result_empty(self.n_rows, np.int32)
        L.check(self._lib.mdb_system_solid_liquid(self._h, int(q6index), float(threshold), int(n_bond),
                                                  int(bool(use_voronoi)), int(nnn), float(rc), L.iptr(sl), L.iptr(nb)))
        return sl, nb

    def rdf_counts(self, rc, nbin, type_list=None, ntype=1, streaming=False):
        """Raw pair counts: (ntype, ntype, nbin) with a type list, (nbin,) for the single-species list kernel."""
        t = L.i32(type_list) if type_list is not None else None
        g = np.zeros((ntype, ntype, nbin) if t is not None else (nbin,), np.float64)
        L.check(self._lib.mdb_system_rdf(self._h, L.iptr(t) if t is not None else None, int(ntype), float(rc),
                                         int(nbin), int(bool(streaming)), L.dptr(g)))
        return g

    def ptm(self, structure="fcc-hcp-bcc", rmsd_threshold=0.1, types=None, fetch=True):
        """(output[n_rows, 8], ptm_indices[n_rows, 18]) on the cached sorted list."""
        out = L.result_empty((self.n_rows, 8), np.float64) if fetch else None
        ind = L.result_empty((self.n_rows, 18), np.int32) if fetch else None
        t = L.i32(types) if types is not None else None
        L.check(self._lib.mdb_system_ptm(self._h, structure.encode(), L.iptr(t) if t is not None else None,
                                         float(rmsd_threshold), L.dptr(out) if fetch else None,
                                         L.iptr(ind) if fetch else None))
        return out, ind

    def planar_faults(self, identify_esf: bool = True):
        """FCC planar-fault labels from the last ``ptm`` call on this handle (identify_fcc_planar_faults.cpp:43)."""
        out = L.result_empty(self.n_rows, np.int32)
        L.check(self._lib.mdb_system_planar_faults(self._h, int(bool(identify_esf)), L.iptr(out)))
        return out

    def chill_plus(self, rc: float, fetch=True):
        """CHILL+ labels on the cached cut-off list (chill_plus.cpp:76)."""
        out = L.result_empty(self.n_rows, np.int32) if fetch else None
        L.check(self._lib.mdb_system_chill_plus(self._h, float(rc), L.iptr(out) if fetch else None))
        return out

    def build_bond(self, type_list, cutoff_matrix):
        """(Nbond, 2) pairs i < j whose listed distance is within cutoff_matrix[type_i, type_j] (build_bond.cpp:9)."""
        t, cm = L.i32(type_list), L.f64(np.asarray(cutoff_matrix, float))
        assert t.shape[0] == self.N and cm.ndim == 2 and cm.shape[0] == cm.shape[1]
        n = C.c_int(0)
        L.check(self._lib.mdb_system_build_bond(self._h, L.iptr(t), L.dptr(cm), cm.shape[0], None, C.byref(n)))
        out = np.empty((n.value, 2), np.int32)
        if n.value:
            L.check(self._lib.mdb_system_build_bond(self._h, L.iptr(t), L.dptr(cm), cm.shape[0], L.iptr(out), C.byref(n)))
        return out

    def voronoi_volume(self):
        """(volume, face count, cavity radius) of every atom's Voronoi cell (voronoi.cpp:16-71)."""
        vol = L.result_empty(self.N, np.float64)
        nn = L.result_empty(self.N, np.int32)
        rad = L.result_empty(self.N, np.float64)
        L.check(self._lib.mdb_system_voronoi_volume(self._h, L.dptr(vol), L.iptr(nn), L.dptr(rad)))
        return vol, nn, rad

    def voronoi_neighbor(self, a_face_area_threshold: float = -1.0, r_face_area_threshold: float = -1.0):
        """(verlet_list, distance_list, face_area, neighbor_number) of voronoi.cpp:307-447."""
        M = C.c_int(0)
        L.check(self._lib.mdb_system_voronoi_neighbor(self._h, float(a_face_area_threshold),
                                                      float(r_face_area_threshold), C.byref(M)))
        m = int(M.value)
        verlet = L.result_empty((self.N, m), np.int32)
        dist = L.result_empty((self.N, m), np.float64)
        area = L.result_empty((self.N, m), np.float64)
        nn = L.result_empty(self.N, np.int32)
        L.check(self._lib.mdb_system_voronoi_fetch(self._h, L.iptr(verlet), L.dptr(dist), L.dptr(area), L.iptr(nn)))
        return verlet, dist, area, nn

    def cnp(self, rc: float, fetch=True):
        """Common neighbour parameter on the cached cut-off list (common_neighbor_parameter.cpp:10)."""
        out = L.result_empty(self.n_rows, np.float64) if fetch else None
        L.check(self._lib.mdb_system_cnp(self._h, float(rc), L.dptr(out) if fetch else None))
        return out

    def wcp(self, type_list, ntype: int):
        """Warren-Cowley matrix (ntype, ntype) from the cached list (warren_cowley_parameter.cpp:9)."""
        t = L.i32(type_list)
        assert t.shape[0] == self.N
        out = np.zeros((int(ntype), int(ntype)), np.float64)
        L.check(self._lib.mdb_system_wcp(self._h, L.iptr(t), int(ntype), L.dptr(out)))
        return out

    def average_by_neighbor(self, rc: float, value, include_self=True):
        """Neighbour average of a per-atom value on the cached list (neighbor.cpp:704)."""
        v = L.f64(value)
        assert v.shape[0] == self.N
        out = L.result_empty(self.n_rows, np.float64)
        L.check(self._lib.mdb_system_average_by_neighbor(self._h, float(rc), L.dptr(v), int(bool(include_self)),
                                                         L.dptr(out)))
        return out

    def cluster(self, rc: float, type_list=None, type1=None, type2=None, r=None):
        """(cluster ids [n_rows], cluster count).  One cut-off, or type-pair cut-offs (cluster.cpp:9-150)."""
        out = L.result_empty(self.n_rows, np.int32)
        cnt = C.c_int(0)
        if type1 is not None:
            t, a, b, rr = L.i32(type_list), L.i32(type1), L.i32(type2), L.f64(r)
            assert t.shape[0] == self.N
            L.check(self._lib.mdb_system_cluster(self._h, float(rc), L.iptr(t), L.iptr(a), L.iptr(b), L.dptr(rr),
                                                 int(a.shape[0]), L.iptr(out), C.byref(cnt)))
        else:
            L.check(self._lib.mdb_system_cluster(self._h, float(rc), None, None, None, None, 0, L.iptr(out),
                                                 C.byref(cnt)))
        return out, cnt.value

    def structure_entropy(self, rc, sigma, use_local_density, volume, average_rc=0.0):
        """(entropy, entropy_ave or None) on the cached cut-off list (structure_entropy.cpp:11)."""
        ent = L.result_empty(self.n_rows, np.float64)
        ave = L.result_empty(self.n_rows, np.float64) if average_rc > 0 else None
        L.check(self._lib.mdb_system_structure_entropy(self._h, float(rc), float(sigma), int(bool(use_local_density)),
                                                       float(volume), float(average_rc), L.dptr(ent),
                                                       L.dptr(ave) if ave is not None else None))
        return ent, ave

    def atomic_temperature(self, vx, vy, vz, mass, rc):
        """Local atomic temperature [K] (velocities A/ps, masses g/mol; atomic_temperature.cpp:7)."""
        a, b, c, m = L.f64(vx), L.f64(vy), L.f64(vz), L.f64(mass)
        assert a.shape[0] == self.N and m.shape[0] == self.N
        out = L.result_empty(self.n_rows, np.float64)
        L.check(self._lib.mdb_system_atomic_temperature(self._h, L.dptr(a), L.dptr(b), L.dptr(c), L.dptr(m), float(rc),
                                                        L.dptr(out)))
        return out

    def bond_analysis(self, rc: float, nbin: int):
        """(bond_length_distribution, bond_angle_distribution), int32[nbin] each (bond_analysis.cpp:7)."""
        bl, ba = np.zeros(int(nbin), np.int32), np.zeros(int(nbin), np.int32)
        L.check(self._lib.mdb_system_bond_analysis(self._h, float(rc) / nbin, 180.0 / nbin, float(rc), int(nbin),
                                                   L.iptr(bl), L.iptr(ba)))
        return bl, ba

    def adf(self, rc_list, pair_list, type_list, nbin: int):
        """Angular distribution per type triplet, int32[Npair, nbin] (bond_analysis.cpp:120)."""
        rcl, pl, t = L.f64(rc_list), L.i32(pair_list), L.i32(type_list)
        assert rcl.shape == (pl.shape[0], 4) and pl.shape[1] == 3 and t.shape[0] == self.N
        out = np.zeros((pl.shape[0], int(nbin)), np.int32)
        L.check(self._lib.mdb_system_adf(self._h, 180.0 / nbin, L.dptr(rcl), L.iptr(pl), pl.shape[0], L.iptr(t), int(nbin),
                                         L.iptr(out)))
        return out

    def result_device(self):
        """Raw device pointers (int) of the latest int32 / f64 per-atom result."""
        a, b = C.c_void_p(), C.c_void_p()
        L.check(self._lib.mdb_system_result_device(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    # -- timing --------------------------------------------------------------
    def set_profiling(self, on=True):
        L.check(self._lib.mdb_system_set_profiling(self._h, int(bool(on))))

    def last_times(self):
        a, b, c = C.c_float(0), C.c_float(0), C.c_float(0)
        L.check(self._lib.mdb_system_last_times(self._h, C.byref(a), C.byref(b), C.byref(c)))
        return {"binning_ms": a.value, "neighbor_ms": b.value, "cna_ms": c.value}


class DeviceGroup:
    """Several GPUs driven by ONE process (section C of include/mdapy_b200.h, csrc/group.cu): the neighbour
    search + CNA path of one unpartitioned host frame, sharded into x slabs with peer stores over NVLink.
    ``devices`` may name a GPU more than once (several slabs on one GPU)."""

    def __init__(self, devices):
        self._lib = L.lib()
        devs = np.ascontiguousarray(np.asarray(list(devices), np.int32).reshape(-1))
        if devs.size < 1:
            raise ValueError("devices must name at least one GPU")
        h = L.c_vp()
        L.check(self._lib.mdb_group_create(L.iptr(devs), int(devs.size), C.byref(h)))
        self._h = h
        self.devices = [int(d) for d in devs]
        self.N = 0
        self.members_used = 0
        self._keep = None
        self.lock = threading.Lock()     # one frame at a time (System objects of several host threads share a group)

    def close(self):
        if getattr(self, "_h", None):
            self._lib.mdb_group_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_atoms(self, x, y, z, box, origin, boundary):
        """Start the upload of a host frame (pageable or page-locked float64 arrays, original atom order)."""
        x, y, z = L.f64(x), L.f64(y), L.f64(z)
        b, o, p = L.box_args(box, origin, boundary)
        L.check(self._lib.mdb_group_set_atoms(self._h, L.dptr(x), L.dptr(y), L.dptr(z), x.shape[0],
                                              L.dptr(b), L.dptr(o), L.iptr(p)))
        self.N = x.shape[0]
        self._keep = (x, y, z)

    def fused_cna(self, rc: float):
        """FixedCNA labels (cna.cpp:429-506) of the uploaded frame in the original atom order."""
        out = L.result_empty(self.N, np.int32)
        used = C.c_int(0)
        L.check(self._lib.mdb_group_fused_cna(self._h, float(rc), L.iptr(out), C.byref(used)))
        self.members_used = int(used.value)
        return out

    def last_times(self):
        t = (C.c_float * 6)()
        L.check(self._lib.mdb_group_last_times(self._h, t))
        keys = ("upload_issue", "route", "compute", "label_push", "download", "total")
        return {k: float(v) for k, v in zip(keys, t)}

    def member_atoms(self):
        out = []
        for d in range(len(self.devices)):
            a, b = C.c_int(0), C.c_int(0)
            L.check(self._lib.mdb_group_member_atoms(self._h, d, C.byref(a), C.byref(b)))
            out.append((int(a.value), int(b.value)))
        return out


_GROUPS = {}
_GROUPS_LOCK = threading.Lock()


def shared_group(devices) -> DeviceGroup:
    """The process-wide DeviceGroup of a device list (System creates one System per frame; the group's
    streams, peer mappings and slab buffers are reused from frame to frame)."""
    key = tuple(int(d) for d in devices)
    with _GROUPS_LOCK:
        g = _GROUPS.get(key)
        if g is None or not getattr(g, "_h", None):
            g = _GROUPS[key] = DeviceGroup(key)
        return g
