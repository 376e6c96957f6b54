"""Host helpers mirroring src/mdapy/tool_function.py (replicate 141-177, _replicate_pos 180-192,
sort_neighbor 75-119).  Replication only ever touches tiny boxes, so it is plain NumPy with the
reference's operation order (repeat_cell.cpp:41-59: shift = ix*a1 + iy*a2 + iz*a3, new = old + shift)."""
from __future__ import annotations

from typing import Tuple

import numpy as np

from . import _lib as L
from .box import Box
from .frame import Frame


def repeat_cell(old_box: np.ndarray, old_pos: np.ndarray, nx: int, ny: int, nz: int) -> np.ndarray:
    a1, a2, a3 = (np.asarray(old_box, float)[k] for k in range(3))
    ix, iy, iz = np.meshgrid(np.arange(nx), np.arange(ny), np.arange(nz), indexing="ij")
    ix, iy, iz = (v.ravel().astype(np.float64)[:, None] for v in (ix, iy, iz))
    shift = ix * a1[None, :] + iy * a2[None, :] + iz * a3[None, :]
    new = np.asarray(old_pos, float)[None, :, :] + shift[:, None, :]
    return np.ascontiguousarray(new.reshape(-1, 3))


def replicate(data: Frame, box: Box, nx: int, ny: int, nz: int) -> Tuple[Frame, Box]:
    nx, ny, nz = int(nx), int(ny), int(nz)
    old_pos = data.select("x", "y", "z").to_numpy()
    new_pos = repeat_cell(box.box, old_pos, nx, ny, nz)
    new_box = box.box * np.array([nx, ny, nz]).reshape((3, 1))
    reps = nx * ny * nz
    cols = {k: np.tile(np.asarray(data[k]), reps) for k in data.columns}
    cols["x"], cols["y"], cols["z"] = new_pos[:, 0].copy(), new_pos[:, 1].copy(), new_pos[:, 2].copy()
    if "id" in cols:
        cols["id"] = np.arange(1, new_pos.shape[0] + 1)
    return Frame(cols), Box(new_box, box.boundary, box.origin)


def _replicate_pos(data: Frame, box: Box, nx: int, ny: int, nz: int) -> Tuple[Frame, Box]:
    nx, ny, nz = int(nx), int(ny), int(nz)
    old_pos = data.select("x", "y", "z").to_numpy()
    new_pos = repeat_cell(box.box, old_pos, nx, ny, nz)
    new_box = box.box * np.array([nx, ny, nz]).reshape((3, 1))
    return (
        Frame({"x": new_pos[:, 0].copy(), "y": new_pos[:, 1].copy(), "z": new_pos[:, 2].copy()}),
        Box(new_box, box.boundary, box.origin),
    )


def sort_neighbor(verlet_list: np.ndarray, distance_list: np.ndarray, neighbor_number: np.ndarray, k: int):
    """In-place partial selection sort of host arrays on the GPU (neighbor.cpp:745-778)."""
    min_number = neighbor_number.min()
    assert min_number >= k, f"The min neighbor number {min_number} is lower than k {k}."
    assert verlet_list.flags.c_contiguous and distance_list.flags.c_contiguous
    N, M = verlet_list.shape
    L.check(L.lib().mdb_sort_verlet_by_distance(L.iptr(verlet_list), L.dptr(distance_list), N, M, int(k), 1))


def average_by_neighbor(average_rc: float, data: Frame, property_name: str, verlet_list: np.ndarray,
                        distance_list: np.ndarray, neighbor_number: np.ndarray, include_self: bool = True,
                        output_name=None) -> Frame:
    """tool_function.py:14-72: neighbour average of a column from HOST lists (one call through the C ABI)."""
    assert property_name in data.columns, f"{property_name} not in data."
    v, d, n = L.i32(verlet_list), L.f64(distance_list), L.i32(neighbor_number)
    value = L.f64(data[property_name])
    out = np.zeros(data.shape[0])
    L.check(L.lib().mdb_average_by_neighbor(float(average_rc), L.iptr(v), v.shape[0], v.shape[1], L.dptr(d),
                                            L.iptr(n), L.dptr(value), L.dptr(out), int(bool(include_self)), 1))
    name = output_name if output_name is not None else f"{property_name}_ave"
    return data.with_columns(**{name: out})


def wrap_pos(data: Frame, box: Box) -> Frame:
    """tool_function.py:122-138: wrap positions into the box along the periodic axes (neighbor.cpp:675)."""
    x, y, z = (np.array(data[c], dtype=np.float64, copy=True) for c in ("x", "y", "z"))
    b, o, p = L.box_args(box.box, box.origin, box.boundary)
    L.check(L.lib().mdb_wrap_positions(L.dptr(x), L.dptr(y), L.dptr(z), x.shape[0], L.dptr(b), L.dptr(o), L.iptr(p), 1))
    return data.with_columns(x=x, y=y, z=z)
