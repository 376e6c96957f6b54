"""ctypes binding of libmdapy_b200.so (the C ABI in include/mdapy_b200.h).

The library is the product: there is no CPU fallback.  Loading fails loudly
when the shared object is missing, and every call fails loudly when no CUDA
device is present (``mdb_system_create`` returns MDB_ERR_CUDA).
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

_PKG = Path(__file__).resolve().parent
LIB_PATH = _PKG / "libmdapy_b200.so"

MDB_OK, MDB_ERR_CUDA, MDB_ERR_VALUE, MDB_ERR_BOX, MDB_ERR_STATE = 0, 1, 2, 3, 4

c_dp = C.POINTER(C.c_double)
c_ip = C.POINTER(C.c_int)
c_fp = C.POINTER(C.c_float)
c_vp = C.c_void_p

_lib = None

# name -> (restype, argtypes); mirrors include/mdapy_b200.h one to one
_BOX = [c_dp, c_dp, c_ip]
_XYZN = [c_dp, c_dp, c_dp, C.c_int]
PROTOTYPES = {
    "mdb_last_error": (C.c_char_p, []),
    "mdb_version": (C.c_char_p, []),
    "mdb_launch_count": (C.c_longlong, []),
    "mdb_device_count": (C.c_int, [c_ip]),
    "mdb_trim_cache": (C.c_int, []),
    "mdb_host_alloc": (C.c_int, [C.c_size_t, C.POINTER(c_vp)]),
    "mdb_host_free": (None, [c_vp]),
    "mdb_build_neighbor": (C.c_int, _XYZN + _BOX + [C.c_double, c_ip, c_dp, c_ip, C.c_int, C.c_int]),
    "mdb_build_neighbor_without_max_neigh": (
        C.c_int, _XYZN + _BOX + [C.c_double, C.c_int, C.POINTER(c_vp), c_ip]),
    "mdb_neighbor_auto_fetch": (C.c_int, [c_vp, c_ip, c_dp, c_ip]),
    "mdb_knn": (C.c_int, _XYZN + _BOX + [C.c_int, c_ip, c_dp, C.c_int]),
    "mdb_sort_verlet_by_distance": (C.c_int, [c_ip, c_dp, C.c_int, C.c_int, C.c_int, C.c_int]),
    "mdb_fcna": (C.c_int, _XYZN + _BOX + [c_ip, C.c_int, c_ip, c_ip, C.c_double, C.c_int]),
    "mdb_acna": (C.c_int, _XYZN + _BOX + [c_ip, C.c_int, c_ip, C.c_int]),
    "mdb_ids": (C.c_int, _XYZN + _BOX + [c_ip, C.c_int, c_ip, c_ip, C.c_int]),
    "mdb_get_csp": (C.c_int, _XYZN + _BOX + [c_ip, C.c_int, C.c_int, c_dp, C.c_int]),
    "mdb_compute_aja": (C.c_int, _XYZN + _BOX + [c_ip, C.c_int, c_dp, C.c_int, c_ip, C.c_int]),
    "mdb_get_ptm": (C.c_int, [C.c_char_p] + _XYZN + _BOX + [c_ip, C.c_int, c_ip, C.c_int, C.c_double, c_dp, C.c_int, c_ip,
                                             C.c_int, C.c_int]),
    "mdb_get_sq": (C.c_int, _XYZN + _BOX + [c_ip, C.c_int, c_dp, c_ip, c_dp, c_ip] + [C.c_int] * 7 +
                   [C.c_double, C.c_int, c_dp, c_dp, c_dp, C.c_int, C.c_int]),
    "mdb_identify_solid_liquid": (C.c_int, [C.c_int, c_dp, c_ip, C.c_int, C.c_int, c_dp, c_ip, c_dp, c_dp, C.c_int,
                                            C.c_int, C.c_double, C.c_int, c_ip, c_ip, C.c_int, C.c_int, C.c_double,
                                            C.c_int]),
    "mdb_rdf": (C.c_int, [c_ip, C.c_int, C.c_int, c_dp, c_ip, c_ip, c_dp, C.c_int, C.c_double, C.c_int]),
    "mdb_rdf_single_species": (C.c_int, [c_ip, C.c_int, C.c_int, c_dp, c_ip, c_dp, C.c_double, C.c_int]),
    "mdb_rdf_streaming": (C.c_int, _XYZN + [c_ip] + _BOX + [c_dp, C.c_int, C.c_double, C.c_int, C.c_int]),
    "mdb_wrap_positions": (C.c_int, [c_dp, c_dp, c_dp, C.c_int] + _BOX + [C.c_int]),
    "mdb_compute_cnp": (C.c_int, _XYZN + _BOX + [c_ip, C.c_int, c_dp, c_ip, c_dp, C.c_double, C.c_int]),
    "mdb_get_wcp": (C.c_int, [c_ip, C.c_int, C.c_int, c_ip, c_ip, C.c_int, c_dp, C.c_int]),
    "mdb_average_by_neighbor": (C.c_int, [C.c_double, c_ip, C.c_int, C.c_int, c_dp, c_ip, c_dp, c_dp, C.c_int, C.c_int]),
    "mdb_get_cluster": (C.c_int, [c_ip, C.c_int, C.c_int, c_dp, c_ip, C.c_double, c_ip, c_ip]),
    "mdb_get_cluster_by_bond": (C.c_int, [c_ip, C.c_int, C.c_int, c_ip, c_ip, c_ip]),
    "mdb_filter_by_type": (C.c_int, [c_ip, C.c_int, C.c_int, c_dp, c_ip, c_ip, c_ip, c_ip, c_dp, C.c_int, C.c_int]),
    "mdb_calculate_structure_entropy": (C.c_int, [C.c_double, C.c_double, C.c_int, C.c_double, c_dp, C.c_int, C.c_int,
                                                  c_ip, c_dp, C.c_int]),
    "mdb_compute_temp": (C.c_int, [c_ip, C.c_int, C.c_int, c_dp, c_dp, c_dp, c_dp, c_dp, c_dp, C.c_double, C.c_int]),
    "mdb_compute_bond": (C.c_int, _XYZN + _BOX + [c_ip, C.c_int, c_dp, c_ip, c_ip, c_ip, C.c_double, C.c_double, C.c_double,
                                                  C.c_int, C.c_int]),
    "mdb_compute_adf": (C.c_int, _XYZN + _BOX + [c_ip, C.c_int, c_dp, c_ip, C.c_double, c_dp, c_ip, C.c_int, c_ip, C.c_int,
                                                 c_ip, C.c_int]),
    "mdb_system_create": (C.c_int, [C.c_int, C.POINTER(c_vp)]),
    "mdb_system_destroy": (None, [c_vp]),
    "mdb_system_set_stream": (C.c_int, [c_vp, c_vp]),
    "mdb_system_synchronize": (C.c_int, [c_vp]),
    "mdb_system_set_atoms": (C.c_int, [c_vp] + _XYZN + _BOX),
    "mdb_system_set_atoms_device": (C.c_int, [c_vp, c_vp, c_vp, c_vp, C.c_int] + _BOX),
    "mdb_system_set_slab_device": (C.c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, C.c_int, C.c_int, C.c_int, C.c_int] + _BOX),
    "mdb_system_set_local_fraction": (C.c_int, [c_vp, C.c_double]),
    "mdb_cell_grid": (C.c_int, _BOX + [C.c_double, c_ip]),
    "mdb_cell_planes_device": (C.c_int, [c_vp, c_vp, c_vp, C.c_int] + _BOX + [C.c_double, c_vp, c_vp]),
    "mdb_slab_pack_device": (C.c_int, [c_vp, c_vp, c_vp, c_vp, C.c_int] + _BOX + [C.c_double, C.c_int, C.c_int, C.c_int,
                                                                                 c_vp, c_vp, C.c_int, c_vp, c_vp]),
    "mdb_slab_unpack_device": (C.c_int, [c_vp, c_vp, C.c_int, c_vp, c_vp, c_vp, c_vp, C.c_int, C.c_int, c_vp, c_vp]),
    "mdb_system_build_neighbor": (C.c_int, [c_vp, C.c_double, C.c_int, c_ip, c_ip]),
    "mdb_system_build_knn": (C.c_int, [c_vp, C.c_int]),
    "mdb_system_sort_neighbor": (C.c_int, [c_vp, C.c_int]),
    "mdb_system_neighbor_min_count": (C.c_int, [c_vp, c_ip]),
    "mdb_system_fetch_neighbor": (C.c_int, [c_vp, c_ip, c_dp, c_ip]),
    "mdb_system_put_neighbor": (C.c_int, [c_vp, c_ip, c_dp, c_ip, C.c_int, C.c_double, C.c_int]),
    "mdb_system_neighbor_device": (C.c_int, [c_vp, C.POINTER(c_vp), C.POINTER(c_vp), C.POINTER(c_vp), c_ip]),
    "mdb_system_fcna": (C.c_int, [c_vp, C.c_double, c_ip]),
    "mdb_system_fused_cna": (C.c_int, [c_vp, C.c_double, c_ip, c_ip]),
    "mdb_identify_sftb_fcc": (C.c_int, [c_ip, C.c_int, c_ip, c_ip, c_ip, C.c_int, c_ip, C.c_int, C.c_int, C.c_int]),
    "mdb_system_planar_faults": (C.c_int, [c_vp, C.c_int, c_ip]),
    "mdb_compute_chill_plus": (C.c_int, _XYZN + _BOX + [c_ip, C.c_int, c_dp, c_ip, C.c_double, c_ip, C.c_int]),
    "mdb_build_bond": (C.c_int, [c_ip, C.c_int, C.c_int, c_dp, c_ip, c_ip, c_dp, C.c_int, c_ip, c_ip, C.c_int]),
    "mdb_system_chill_plus": (C.c_int, [c_vp, C.c_double, c_ip]),
    "mdb_system_build_bond": (C.c_int, [c_vp, c_ip, c_dp, C.c_int, c_ip, c_ip]),
    "mdb_repeat_cell": (C.c_int, [c_dp, c_dp, c_dp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]),
    "mdb_transform_and_filter": (C.c_int, _XYZN + [c_dp, c_dp, c_dp, c_dp, C.c_int, c_dp, c_ip, C.c_int]),
    "mdb_filter_overlap_atom": (C.c_int, _XYZN + _BOX + [C.c_double, c_vp, C.c_int]),
    "mdb_system_set_atoms_lattice": (C.c_int, [c_vp, c_dp, c_dp, C.c_int, C.c_int, C.c_int, C.c_int, c_dp, c_ip]),
    "mdb_system_positions_device": (C.c_int, [c_vp, C.POINTER(c_vp), C.POINTER(c_vp), C.POINTER(c_vp), c_ip]),
    "mdb_system_fetch_positions": (C.c_int, [c_vp, c_dp, c_dp, c_dp]),
    "mdb_system_filter_overlap": (C.c_int, [c_vp, C.c_double, c_vp]),
    "mdb_system_transform_and_filter": (C.c_int, [c_vp, c_dp, c_dp, c_dp, c_dp, C.c_int, c_dp, c_ip]),
    "mdb_system_acna": (C.c_int, [c_vp, c_ip]),
    "mdb_system_ids": (C.c_int, [c_vp, c_ip]),
    "mdb_system_csp": (C.c_int, [c_vp, C.c_int, c_dp]),
    "mdb_system_aja": (C.c_int, [c_vp, c_ip]),
    "mdb_system_steinhardt": (C.c_int, [c_vp, c_ip, C.c_int, C.c_int, C.c_double, C.c_int, C.c_int, C.c_int, C.c_int,
                                        c_dp, c_dp, c_dp, c_dp]),
    "mdb_system_solid_liquid": (C.c_int, [c_vp, C.c_int, C.c_double, C.c_int, C.c_int, C.c_int, C.c_double, c_ip,
                                          c_ip]),
    "mdb_system_rdf": (C.c_int, [c_vp, c_ip, C.c_int, C.c_double, C.c_int, C.c_int, c_dp]),
    "mdb_system_ptm": (C.c_int, [c_vp, C.c_char_p, c_ip, C.c_double, c_dp, c_ip]),
    "mdb_system_cnp": (C.c_int, [c_vp, C.c_double, c_dp]),
    "mdb_system_wcp": (C.c_int, [c_vp, c_ip, C.c_int, c_dp]),
    "mdb_system_average_by_neighbor": (C.c_int, [c_vp, C.c_double, c_dp, C.c_int, c_dp]),
    "mdb_system_cluster": (C.c_int, [c_vp, C.c_double, c_ip, c_ip, c_ip, c_dp, C.c_int, c_ip, c_ip]),
    "mdb_system_structure_entropy": (C.c_int, [c_vp, C.c_double, C.c_double, C.c_int, C.c_double, C.c_double, c_dp,
                                               c_dp]),
    "mdb_system_check_small_division": (C.c_int, [c_vp, c_dp, C.c_int, C.c_int, C.POINTER(C.c_longlong)]),
    "mdb_system_atomic_temperature": (C.c_int, [c_vp, c_dp, c_dp, c_dp, c_dp, C.c_double, c_dp]),
    "mdb_system_bond_analysis": (C.c_int, [c_vp, C.c_double, C.c_double, C.c_double, C.c_int, c_ip, c_ip]),
    "mdb_system_adf": (C.c_int, [c_vp, C.c_double, c_dp, c_ip, C.c_int, c_ip, C.c_int, c_ip]),
    "mdb_system_result_device": (C.c_int, [c_vp, C.POINTER(c_vp), C.POINTER(c_vp)]),
    "mdb_system_set_profiling": (C.c_int, [c_vp, C.c_int]),
    "mdb_system_last_times": (C.c_int, [c_vp, c_fp, c_fp, c_fp]),
    "mdb_get_voronoi_volume_number_radius": (C.c_int, _XYZN + _BOX + [c_dp, c_ip, c_dp, C.c_int]),
    "mdb_system_voronoi_volume": (C.c_int, [c_vp, c_dp, c_ip, c_dp]),
    "mdb_system_voronoi_neighbor": (C.c_int, [c_vp, C.c_double, C.c_double, c_ip]),
    "mdb_system_voronoi_fetch": (C.c_int, [c_vp, c_ip, c_dp, c_dp, c_ip]),
    # section C: device group (one process, several GPUs)
    "mdb_group_create": (C.c_int, [c_ip, C.c_int, C.POINTER(c_vp)]),
    "mdb_group_destroy": (None, [c_vp]),
    "mdb_group_size": (C.c_int, [c_vp]),
    "mdb_group_set_atoms": (C.c_int, [c_vp, c_dp, c_dp, c_dp, C.c_int] + _BOX),
    "mdb_group_fused_cna": (C.c_int, [c_vp, C.c_double, c_ip, c_ip]),
    "mdb_group_last_times": (C.c_int, [c_vp, c_fp]),
    "mdb_group_member_atoms": (C.c_int, [c_vp, C.c_int, c_ip, c_ip]),
}


def lib() -> C.CDLL:
    """Load libmdapy_b200.so (built by ``make -C mdapy_b200/csrc`` / __graft_entry__.build())."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise ImportError(
                f"{LIB_PATH} not found: the CUDA extension is the only backend of mdapy_b200 "
                "(no CPU fallback). Build it with `python -c 'import __graft_entry__ as g; g.build()'`."
            )
        L = C.CDLL(str(LIB_PATH))
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(L, name)  # AttributeError here = header/library mismatch
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(code: int) -> None:
    """Map MDB_ERR_* to the exception types the reference raises at this boundary."""
    if code == MDB_OK:
        return
    msg = lib().mdb_last_error().decode(errors="replace")
    if code == MDB_ERR_VALUE:
        raise ValueError(msg)
    raise RuntimeError(msg)


def dptr(a: np.ndarray):
    return a.ctypes.data_as(c_dp)


def iptr(a: np.ndarray):
    return a.ctypes.data_as(c_ip)


def f64(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float64)


def i32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.int32)


def box_args(box, origin, boundary):
    """(3,3) f64 box, (3,) f64 origin, (3,) int32 boundary -- int64 flags are converted
    like nanobind's implicit cast does for the reference (box.py:247-259)."""
    b = f64(box).reshape(3, 3)
    o = f64(origin).reshape(3)
    p = i32(boundary).reshape(3)
    return b, o, p


class _PinnedBlock:
    """Page-locked host block from the library's cache, exposed through the array interface."""

    def __init__(self, nbytes: int, shape, dtype):
        ptr = c_vp()
        check(lib().mdb_host_alloc(nbytes, C.byref(ptr)))
        self._ptr = ptr.value
        self.__array_interface__ = {"shape": tuple(shape), "typestr": np.dtype(dtype).str,
                                    "data": (self._ptr, False), "version": 3}

    def __del__(self):
        if getattr(self, "_ptr", None) and _lib is not None:
            _lib.mdb_host_free(c_vp(self._ptr))
            self._ptr = None


PINNED_MIN_BYTES = 1 << 20
PINNED_MAX_BYTES = 4 << 30


def result_empty(shape, dtype) -> np.ndarray:
    """Uninitialised host array for a result column.  Between 1 MiB and 4 GiB it lives in page-locked
    memory from the library's cache, so the device -> host copy runs at PCIe rate without a staging
    pass and without first-touch page faults; the block returns to the cache when the array dies."""
    shape = (shape,) if np.isscalar(shape) else tuple(shape)
    nbytes = int(np.prod(shape, dtype=np.int64)) * np.dtype(dtype).itemsize
    if nbytes < PINNED_MIN_BYTES or nbytes > PINNED_MAX_BYTES:
        return np.empty(shape, dtype)
    return np.asarray(_PinnedBlock(nbytes, shape, dtype))


def empty_cache() -> None:
    """Return the cached device and pinned-host blocks to the driver."""
    check(lib().mdb_trim_cache())
