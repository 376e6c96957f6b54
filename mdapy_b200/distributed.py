"""Spatial domain decomposition of one frame across the GPUs of a box.

One process per GPU (``torch.distributed``, NCCL).  The GLOBAL cell grid of the cut-off search
(``ncell = max(floor(thickness/rc), 3)``, src/neighbor.cpp:367-370) is cut into contiguous slabs of
x cell planes; rank r owns the atoms whose global x plane lies in ``[lo_r, hi_r)``.  The only data
exchange of a neighbour build is one halo step: every rank sends the atoms of its first owned plane
to its left neighbour and those of its last owned plane to its right neighbour (periodic ring,
exactly the planes the reference's 27-cell stencil with ``mod()`` wrap touches), as raw global
coordinates + global id.  Each rank then runs the ordinary kernels on owned + ghost atoms with the
GLOBAL box, so distances, membership AND row order are bit-identical to the single-GPU build
(cells are global, ties inside a cell are ordered by global id).

Not decomposed (round 1): kNN-based descriptors (adaptive halo), Steinhardt averaging / solid-liquid
(2*rc halo).
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np

from . import _lib as L


def cell_grid(box, origin, boundary, rc):
    """Cell counts of the global cut-off grid (host arithmetic of the library, no GPU needed)."""
    b, o, p = L.box_args(box, origin, boundary)
    n = (C.c_int * 3)()
    L.check(L.lib().mdb_cell_grid(L.dptr(b), L.dptr(o), L.iptr(p), float(rc), n))
    return [int(n[0]), int(n[1]), int(n[2])]


def slab_bounds(n0: int, world: int):
    """Plane ranges [lo_r, hi_r): as even as integer division allows."""
    if n0 < 3 * world:
        raise ValueError(
            f"cell grid has {n0} x-planes; a decomposition over {world} ranks needs >= {3 * world} "
            "(owned slab + two ghost planes must not overlap around the ring)")
    return [r * n0 // world for r in range(world + 1)]


class SlabDecomposition:
    def __init__(self, box, origin, boundary, rc, rank: int, world: int, device=None, group=None):
        import torch

        self.torch = torch
        self.box = np.ascontiguousarray(np.asarray(box, float).reshape(3, 3))
        self.origin = np.asarray(origin, float).reshape(3)
        self.boundary = np.asarray(boundary, np.int32).reshape(3)
        self.rc = float(rc)
        self.rank, self.world = int(rank), int(world)
        self.device = device
        self.group = group
        self.ncell = cell_grid(self.box, self.origin, self.boundary, self.rc)
        self.n0 = self.ncell[0]
        self.bounds = slab_bounds(self.n0, self.world)
        self.lo, self.hi = self.bounds[self.rank], self.bounds[self.rank + 1]
        self.left = (self.rank - 1) % self.world
        self.right = (self.rank + 1) % self.world
        self.plane0 = (self.lo - 1) % self.n0          # first stored plane (left ghost)
        self.nplanes = self.hi - self.lo + 2
        self._bounds_t = None
        self._ds = None
        self.halo_atoms = 0

    # ------------------------------------------------------------------ plane bookkeeping
    def planes(self, x, y, z):
        """Global x cell plane of every atom (device kernel, reference cell arithmetic)."""
        torch = self.torch
        out = torch.empty(x.shape[0], dtype=torch.int32, device=x.device)
        b, o, p = L.box_args(self.box, self.origin, self.boundary)
        stream = torch.cuda.current_stream().cuda_stream
        L.check(L.lib().mdb_cell_planes_device(
            C.c_void_p(x.data_ptr()), C.c_void_p(y.data_ptr()), C.c_void_p(z.data_ptr()), int(x.shape[0]),
            L.dptr(b), L.dptr(o), L.iptr(p), self.rc, C.c_void_p(out.data_ptr()), C.c_void_p(int(stream))))
        return out

    def owner_of(self, planes):
        torch = self.torch
        if self._bounds_t is None or self._bounds_t.device != planes.device:
            self._bounds_t = torch.tensor(self.bounds[1:-1], dtype=planes.dtype, device=planes.device)
        return torch.bucketize(planes, self._bounds_t, right=True)

    def lattice_planes(self, n: int, a: float):
        """Lattice x-plane range [ix0, ix1) that surely covers this rank's slab of an n-cell cubic crystal."""
        x_lo = self.lo * self.rc
        x_hi = n * a if self.hi == self.n0 else self.hi * self.rc
        ix0 = max(int(np.floor(x_lo / a)) - 1, 0)
        ix1 = min(int(np.ceil(x_hi / a)) + 1, n)
        return ix0, ix1

    # ------------------------------------------------------------------ collectives
    def _all_to_all(self, packed, send_counts):
        """Variable all-to-all of rows of ``packed`` (already grouped by destination rank)."""
        torch = self.torch
        import torch.distributed as dist

        sc = torch.tensor(send_counts, dtype=torch.int64, device=packed.device)
        rcnt = torch.empty_like(sc)
        dist.all_to_all_single(rcnt, sc, group=self.group)
        rc_list = [int(v) for v in rcnt.tolist()]
        out = torch.empty((sum(rc_list), packed.shape[1]), dtype=packed.dtype, device=packed.device)
        dist.all_to_all_single(out, packed, output_split_sizes=rc_list, input_split_sizes=list(send_counts),
                               group=self.group)
        return out

    @staticmethod
    def _pack(torch, x, y, z, gid, sel):
        return torch.stack([x[sel], y[sel], z[sel], gid[sel].to(torch.float64)], dim=1)

    def migrate(self, x, y, z, gid, planes=None):
        """Send every atom to the rank that owns its cell plane (input distribution step)."""
        torch = self.torch
        if planes is None:
            planes = self.planes(x, y, z)
        owner = self.owner_of(planes)
        order = torch.argsort(owner, stable=True)
        counts = torch.bincount(owner, minlength=self.world).tolist()
        packed = self._pack(torch, x, y, z, gid, order)
        got = self._all_to_all(packed, [int(c) for c in counts])
        return (got[:, 0].contiguous(), got[:, 1].contiguous(), got[:, 2].contiguous(),
                got[:, 3].to(torch.int32).contiguous())

    def exchange_halo(self, x, y, z, gid, planes):
        """Ghost atoms of this rank: the neighbours' boundary planes (raw coordinates, global ids)."""
        torch = self.torch
        first = torch.nonzero(planes == self.lo).flatten()
        last = torch.nonzero(planes == self.hi - 1).flatten()
        counts = [0] * self.world
        if self.left == self.right:      # two ranks: both boundary planes go to the same peer
            sel = torch.cat([first, last])
            counts[self.left] = int(sel.numel())
        else:
            lo_first = self.left < self.right
            sel = torch.cat([first, last]) if lo_first else torch.cat([last, first])
            counts[self.left] = int(first.numel())
            counts[self.right] = int(last.numel())
        got = self._all_to_all(self._pack(torch, x, y, z, gid, sel), counts)
        self.halo_atoms = int(got.shape[0])
        return (got[:, 0].contiguous(), got[:, 1].contiguous(), got[:, 2].contiguous(),
                got[:, 3].to(torch.int32).contiguous())

    # ------------------------------------------------------------------ one frame
    def build(self, x, y, z, gid, max_neigh=None, device_index: Optional[int] = None):
        """Halo exchange + local cut-off neighbour build for the owned atoms ``x, y, z, gid``
        (torch CUDA tensors).  Returns the DeviceSystem holding the slab lists (n_owned rows)."""
        torch = self.torch
        from .device import DeviceSystem

        planes = self.planes(x, y, z)
        gx, gy, gz, gg = self.exchange_halo(x, y, z, gid, planes)
        ax, ay, az = torch.cat([x, gx]), torch.cat([y, gy]), torch.cat([z, gz])
        ag = torch.cat([gid, gg])
        if self._ds is None:
            idx = device_index if device_index is not None else torch.cuda.current_device()
            self._ds = DeviceSystem(idx)
        ds = self._ds
        ds.set_slab_device(ax, ay, az, ag, int(x.shape[0]), self.plane0, self.nplanes, self.box, self.origin,
                           self.boundary, stream=torch.cuda.current_stream().cuda_stream)
        ds.build_neighbor(self.rc, max_neigh)
        return ds

    def make_step(self, x, y, z, gid):
        """Benchmark step: one frame = planes + halo exchange + binning + neighbour build + CNA."""

        def step():
            ds = self.build(x, y, z, gid)
            ds.fcna(self.rc, fetch=False)
            return ds.M

        return step
