"""Input side of the hot path on the device: the three C++ helpers the reference's crystal /
polycrystal builders call, with the same argument meaning, plus a device-resident lattice generator.

    repeat_cell(old_box, old_pos, nx, ny, nz)                         <- _repeat_cell.repeat_cell   (repeat_cell.cpp:19)
    transform_and_filter(x, y, z, rotation, center, target, coeffs)   <- _polycrystal.transform_and_filter (polycrystal.cpp:21)
    filter_overlap_atom(x, y, z, box, origin, boundary, rc)           <- _neighbor.filter_overlap_atom (neighbor.cpp:390)
    DeviceSystem lattice: device_lattice(structure, a, nx, ny, nz)    build_crystal's frame generated in HBM

Results are bit-identical to the reference's (tests/test_gpu_builders.py)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib as L
from .device import DeviceSystem
from .lattice import _BASES


def repeat_cell(old_box, old_pos, nx: int, ny: int, nz: int) -> np.ndarray:
    b = L.f64(np.asarray(old_box, float).reshape(3, 3))
    p = L.f64(np.asarray(old_pos, float).reshape(-1, 3))
    out = np.empty((p.shape[0] * int(nx) * int(ny) * int(nz), 3), np.float64)
    L.check(L.lib().mdb_repeat_cell(L.dptr(out), L.dptr(b), L.dptr(p), p.shape[0], int(nx), int(ny), int(nz), 0))
    return out


def transform_and_filter(x, y, z, rotation_matrix, center, target_center, coeffs) -> np.ndarray:
    x, y, z = L.f64(x), L.f64(y), L.f64(z)
    R = L.f64(np.asarray(rotation_matrix, float).reshape(3, 3))
    c, t = L.f64(np.asarray(center, float).reshape(3)), L.f64(np.asarray(target_center, float).reshape(3))
    pl = L.f64(np.asarray(coeffs, float).reshape(-1, 4))
    out = np.empty((x.shape[0], 3), np.float64)
    n = C.c_int(0)
    L.check(L.lib().mdb_transform_and_filter(L.dptr(x), L.dptr(y), L.dptr(z), x.shape[0], L.dptr(R), L.dptr(c), L.dptr(t),
                                             L.dptr(pl), pl.shape[0], L.dptr(out), C.byref(n), 0))
    return out[: n.value].copy()


def filter_overlap_atom(x, y, z, box, origin, boundary, rc: float) -> np.ndarray:
    x, y, z = L.f64(x), L.f64(y), L.f64(z)
    b, o, p = L.box_args(box, origin, boundary)
    keep = np.empty(x.shape[0], np.uint8)
    L.check(L.lib().mdb_filter_overlap_atom(L.dptr(x), L.dptr(y), L.dptr(z), x.shape[0], L.dptr(b), L.dptr(o), L.iptr(p),
                                            float(rc), keep.ctypes.data_as(C.c_void_p), 0))
    return keep.astype(bool)


def device_lattice(structure: str, a: float, nx: int, ny: int, nz: int, boundary=(1, 1, 1), device: int = 0) -> DeviceSystem:
    """``build_crystal(structure, a, nx, ny, nz)`` generated straight into a DeviceSystem (positions never
    touch the host; 100 M atoms take ~1 ms instead of a 2.4 GB upload)."""
    s = structure.lower()
    if s not in _BASES:
        raise ValueError(f"structure {structure!r} is not available here; supported: {sorted(_BASES)}")
    cell = L.f64(a * np.eye(3))
    basis = L.f64(_BASES[s] @ (a * np.eye(3)))
    ds = DeviceSystem(device)
    o, p = L.f64(np.zeros(3)), L.i32(np.asarray(boundary, np.int32))
    L.check(ds._lib.mdb_system_set_atoms_lattice(ds._h, L.dptr(cell), L.dptr(basis), basis.shape[0], int(nx), int(ny), int(nz),
                                                 L.dptr(o), L.iptr(p)))
    ds.N = ds.n_rows = basis.shape[0] * int(nx) * int(ny) * int(nz)
    ds.box = cell * np.array([nx, ny, nz], float).reshape(3, 1)
    return ds


def fetch_positions(ds: DeviceSystem):
    x, y, z = (L.result_empty(ds.N, np.float64) for _ in range(3))
    L.check(ds._lib.mdb_system_fetch_positions(ds._h, L.dptr(x), L.dptr(y), L.dptr(z)))
    return x, y, z


def bisector_planes(seeds: np.ndarray, g: int, box_lengths: np.ndarray) -> np.ndarray:
    """Plane coefficients (a, b, c, d) with a x + b y + c z + d < 0 inside the periodic Voronoi cell of seed g:
    the perpendicular bisectors towards every periodic image of every other seed (a superset of the cell's
    faces -- the redundant planes do not change the inside test)."""
    s = seeds[g]
    rows = []
    L3 = np.asarray(box_lengths, float)
    for h in range(seeds.shape[0]):
        for ix in (-1, 0, 1):
            for iy in (-1, 0, 1):
                for iz in (-1, 0, 1):
                    if h == g and ix == 0 and iy == 0 and iz == 0:
                        continue
                    q = seeds[h] + np.array([ix, iy, iz]) * L3
                    n = q - s
                    # |p - s|^2 < |p - q|^2  <=>  2 n.p - (|q|^2 - |s|^2) < 0
                    rows.append([2 * n[0], 2 * n[1], 2 * n[2], -(q @ q - s @ s)])
    return np.asarray(rows, float)
