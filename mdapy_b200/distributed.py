"""Spatial domain decomposition of one frame across the GPUs of a box.

One process per GPU (``torch.distributed``, NCCL).  The GLOBAL cell grid of the cut-off search
(``ncell = max(floor(thickness/rc), 3)``, src/neighbor.cpp:367-370) is cut into contiguous slabs of
x cell planes; rank r owns the atoms whose global x plane lies in ``[lo_r, hi_r)``.  The only data
exchange of a neighbour build is one halo step: every rank sends the atoms of its first ``halo`` owned
planes to its left neighbour and those of its last ``halo`` planes to its right neighbour (periodic
ring, exactly the planes the reference's 27-cell stencil with ``mod()`` wrap touches), as raw global
coordinates + global id (+ optional per-atom payload such as the type).  Each rank then runs the
ordinary kernels on owned + ghost atoms with the GLOBAL box, so distances, membership AND row order
are bit-identical to the single-GPU build (cells are global, ties inside a cell are ordered by global
id).

Halo depth.  ``halo = 1`` serves everything that reads only the COORDINATES of an
atom's neighbours: the cut-off list, fixed-cutoff CNA, CSP / Ackland-Jones from the sorted cut-off
list, Steinhardt q_l / w_l without averaging, RDF.  A descriptor that reads a quantity DERIVED from
the neighbours' own lists needs one more plane per level of indirection: the local frame then keeps
list rows for the owned atoms and the inner ``halo - 1`` ghost layers (``n_rows``), the outermost layer
serves as neighbours only, and results are valid for the owned atoms: Steinhardt averaging 2,
identifySolidLiquid 3 (4 on averaged q_6).

k-nearest lists (CSP, Ackland-Jones, adaptive CNA, PTM, diamond identification without a cached
cut-off list) have no a-priori radius: :class:`KnnDecomposition` cuts the box into planes of a nominal
width derived from the density, exchanges ``halo`` planes, runs the ordinary k-nearest search over
owned + ghost atoms and VERIFIES that no needed row reaches past the halo (k-th distance <= distance
to the outer halo edge); it widens the halo and repeats otherwise.
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


def slab_bounds(n0: int, world: int, halo: int = 1):
    """Plane ranges [lo_r, hi_r): as even as integer division allows."""
    if n0 < (2 * halo + 1) * world and world > 1:
        raise ValueError(
            f"cell grid has {n0} x-planes; a decomposition over {world} ranks with halo {halo} needs >= "
            f"{(2 * halo + 1) * world} (owned slab + ghost planes on both sides must not overlap around the ring)")
    return [r * n0 // world for r in range(world + 1)]


class _DeviceView:
    """Zero-copy view of library-owned device memory for ``torch.as_tensor``."""

    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 3, "strides": None}


class SlabDecomposition:
    def __init__(self, box, origin, boundary, rc, rank: int, world: int, device=None, group=None, halo: int = 1):
        import torch

        self.torch = torch
        self.box = np.ascontiguousarray(np.asarray(box, float).reshape(3, 3))
        self.origin = np.asarray(origin, float).reshape(3)
        self.boundary = np.asarray(boundary, np.int32).reshape(3)
        self.rc = float(rc)
        self.rank, self.world = int(rank), int(world)
        self.device = device
        self.group = group
        self.halo = int(halo)
        assert self.halo >= 1
        self.ncell = cell_grid(self.box, self.origin, self.boundary, self.rc)
        self.n0 = self.ncell[0]
        self.bounds = slab_bounds(self.n0, self.world, self.halo)
        self.lo, self.hi = self.bounds[self.rank], self.bounds[self.rank + 1]
        self.left = (self.rank - 1) % self.world
        self.right = (self.rank + 1) % self.world
        self.plane0 = (self.lo - self.halo) % self.n0          # first stored plane (outermost left ghost)
        self.nplanes = self.hi - self.lo + 2 * self.halo
        self._bounds_t = None
        self._ds = None
        self.halo_atoms = 0
        self.n_owned = 0
        self.n_rows = 0

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

    def ghost_layer(self, planes):
        """1..halo for a plane of the left / right ghost region of this rank, 0 otherwise."""
        torch = self.torch
        dl = torch.remainder(self.lo - planes, self.n0)            # 1..halo on the left
        dr = torch.remainder(planes - (self.hi - 1), self.n0)      # 1..halo on the right
        layer = torch.zeros_like(planes)
        layer = torch.where((dl >= 1) & (dl <= self.halo), dl, layer)
        layer = torch.where((dr >= 1) & (dr <= self.halo), dr, layer)
        return layer

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
    def _pack(torch, x, y, z, gid, sel, extra=None):
        cols = [x[sel], y[sel], z[sel], gid[sel].to(torch.float64)]
        if extra is not None:
            cols += [e[sel].to(torch.float64) for e in extra]
        return torch.stack(cols, dim=1)

    @staticmethod
    def _unpack(torch, got, extra=None):
        out = [got[:, 0].contiguous(), got[:, 1].contiguous(), got[:, 2].contiguous(),
               got[:, 3].to(torch.int32).contiguous()]
        ext = None
        if extra is not None:
            ext = [got[:, 4 + i].to(e.dtype).contiguous() for i, e in enumerate(extra)]
        return out, ext

    def migrate(self, x, y, z, gid, planes=None, extra=None):
        """Send every atom to the rank that owns its cell plane (input distribution step)."""
        torch = self.torch
        if planes is None:
            planes = self.planes(x, y, z)
        owner = self.owner_of(planes)
        order = torch.argsort(owner, stable=True)
        counts = torch.bincount(owner, minlength=self.world).tolist()
        packed = self._pack(torch, x, y, z, gid, order, extra)
        got = self._all_to_all(packed, [int(c) for c in counts])
        (gx, gy, gz, gg), ext = self._unpack(torch, got, extra)
        if extra is None:
            return gx, gy, gz, gg
        return gx, gy, gz, gg, ext

    def exchange_halo(self, x, y, z, gid, planes, extra=None):
        """Ghost atoms of this rank: the neighbours' ``halo`` boundary planes (raw coordinates, global
        ids, optional per-atom payload).  Returns (gx, gy, gz, ggid) [+ payload list]."""
        torch = self.torch
        h = self.halo
        first = torch.nonzero((planes >= self.lo) & (planes < self.lo + h)).flatten()
        last = torch.nonzero((planes >= self.hi - h) & (planes < self.hi)).flatten()
        counts = [0] * self.world
        if self.world == 1:              # one rank: its own periodic images are found by the global wrap
            got = self._pack(torch, x, y, z, gid, first[:0], extra)
            self.halo_atoms = 0
            (gx, gy, gz, gg), ext = self._unpack(torch, got, extra)
            return (gx, gy, gz, gg) if extra is None else (gx, gy, gz, gg, ext)
        if self.left == self.right:    # two ranks: both boundary regions go to the same peer
            sel = torch.cat([first, last])
            counts[self.left] = int(sel.numel())
        else:
            lo_first = self.left < self.right
            sel = torch.cat([first, last]) if lo_first else torch.cat([last, first])
            counts[self.left] = int(first.numel())
            counts[self.right] = int(last.numel())
        got = self._all_to_all(self._pack(torch, x, y, z, gid, sel, extra), counts)
        self.halo_atoms = int(got.shape[0])
        (gx, gy, gz, gg), ext = self._unpack(torch, got, extra)
        if extra is None:
            return gx, gy, gz, gg
        return gx, gy, gz, gg, ext

    # ------------------------------------------------------------------ one frame
    def assemble(self, x, y, z, gid, extra=None):
        """Halo exchange + local ordering [owned | inner ghost layers | outermost ghost layer].
        Returns (ax, ay, az, agid, extras) and sets n_owned / n_rows."""
        torch = self.torch
        planes = self.planes(x, y, z)
        if extra is None:
            gx, gy, gz, gg = self.exchange_halo(x, y, z, gid, planes)
            gext = None
        else:
            gx, gy, gz, gg, gext = self.exchange_halo(x, y, z, gid, planes, extra)
        self.n_owned = int(x.shape[0])
        n_inner = 0
        if self.halo > 1 and gx.numel():
            layer = self.ghost_layer(self.planes(gx, gy, gz))
            order = torch.argsort(layer, stable=True)       # inner layers first, outermost (== halo) last
            gx, gy, gz, gg = gx[order], gy[order], gz[order], gg[order]
            if gext is not None:
                gext = [e[order] for e in gext]
            n_inner = int((layer < self.halo).sum().item())
        self.n_rows = self.n_owned + n_inner
        ax, ay, az = torch.cat([x, gx]), torch.cat([y, gy]), torch.cat([z, gz])
        ag = torch.cat([gid, gg])
        aext = None if extra is None else [torch.cat([e, g]) for e, g in zip(extra, gext)]
        return ax, ay, az, ag, aext

    def select_replicated(self, x, y, z, gid, extra=None):
        """Every rank holds the WHOLE frame (e.g. all ranks read the same file): pick this rank's owned
        and ghost atoms locally, no communication.  Same return value and ordering as :meth:`assemble`."""
        torch = self.torch
        planes = self.planes(x, y, z)
        own = torch.nonzero((planes >= self.lo) & (planes < self.hi)).flatten()
        if self.world == 1:
            gi = own[:0]
            n_inner = 0
        else:
            layer = self.ghost_layer(planes)
            gi = torch.nonzero(layer > 0).flatten()
            gl = layer[gi]
            order = torch.argsort(gl, stable=True)
            gi = gi[order]
            n_inner = int((gl < self.halo).sum().item())
        sel = torch.cat([own, gi])
        self.n_owned = int(own.numel())
        self.n_rows = self.n_owned + n_inner
        self.halo_atoms = int(gi.numel())
        aext = None if extra is None else [e[sel].contiguous() for e in extra]
        return x[sel].contiguous(), y[sel].contiguous(), z[sel].contiguous(), gid[sel].contiguous(), aext

    def device_system(self, device_index: Optional[int] = None):
        from .device import DeviceSystem

        if self._ds is None:
            idx = device_index if device_index is not None else self.torch.cuda.current_device()
            self._ds = DeviceSystem(idx)
        return self._ds

    def load(self, x, y, z, gid, extra=None, device_index: Optional[int] = None, replicated: bool = False):
        """Halo exchange (or local selection from a replicated frame), then hand owned + ghost atoms to
        this rank's DeviceSystem (no list yet)."""
        torch = self.torch
        # the library reads x/y/z as f64 and ids as int32 device arrays: convert once, never reinterpret
        x, y, z = (t.to(device=self.device, dtype=torch.float64).contiguous() for t in (x, y, z))
        if gid.dtype != torch.int32:
            assert int(gid.max()) < 2 ** 31, "global ids must fit int32"
        gid = gid.to(device=self.device, dtype=torch.int32).contiguous()
        if replicated:
            ax, ay, az, ag, aext = self.select_replicated(x, y, z, gid, extra)
        else:
            ax, ay, az, ag, aext = self.assemble(x, y, z, gid, extra)
        ds = self.device_system(device_index)
        stream = torch.cuda.current_stream().cuda_stream
        if self.world == 1:   # nothing to exchange: the ordinary single-GPU frame
            ds.set_atoms_device(ax, ay, az, self.box, self.origin, self.boundary, stream=stream)
        else:
            ds.set_slab_device(ax, ay, az, ag, self.n_rows, self.plane0, self.nplanes, self.box, self.origin,
                               self.boundary, stream=stream)
            ds.set_local_fraction(min(1.0, self.nplanes / self.n0))
        self.local_extra = aext
        self.local = (ax, ay, az, ag)
        return ds

    def build(self, x, y, z, gid, max_neigh=None, device_index: Optional[int] = None, extra=None,
              sync_width: bool = False, replicated: bool = False):
        """Halo exchange + local cut-off neighbour build for the owned atoms ``x, y, z, gid``
        (torch CUDA tensors).  Returns the DeviceSystem holding the slab lists (n_rows rows, the first
        n_owned of them owned).  ``sync_width``: make the automatic row width the global maximum
        (one all-reduce), so exported lists have the single-GPU shape."""
        ds = self.load(x, y, z, gid, extra, device_index, replicated)
        M, mx = ds.build_neighbor(self.rc, max_neigh)
        if sync_width and max_neigh is None and self.world > 1:
            import torch.distributed as dist

            assert dist.is_initialized(), "sync_width needs an initialised process group"
            t = self.torch.tensor([M], dtype=self.torch.int64, device=x.device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)
            if int(t.item()) != M:
                ds.build_neighbor(self.rc, int(t.item()))   # binning is cached: only the fill pass reruns
        return ds

    # ------------------------------------------------------------------ resident frames (halo == 1)
    # Trajectory / benchmark use: the slab lives in ONE preallocated buffer per coordinate, owned atoms at its
    # head.  Per frame the device packs the boundary planes (one kernel), the two fixed-capacity buffers travel
    # over NCCL (all_to_all_single with constant split sizes: no count exchange, no host round trip), and a
    # second kernel appends the received ghosts behind the owned atoms.  The only host read per frame is the new
    # atom count (one int).  PyTorch is the NCCL / memory plumbing; every byte of the data path is moved by the
    # library's kernels.
    def resident_buffers(self, n_owned: int, cap: Optional[int] = None):
        """Allocate the frame buffers for ``n_owned`` owned atoms and return the (x, y, z, gid) views the
        caller fills (device tensors; a host -> device copy can target them directly)."""
        torch = self.torch
        assert self.halo == 1, "the resident fast path covers halo == 1 (cut-off list, CNA, CSP, Steinhardt, RDF)"
        if cap is None:   # rows per send buffer: one plane of the slab, + 50 %; the SAME on every rank
            cap = int(1.5 * n_owned / max(self.hi - self.lo, 1)) + 4096
            import torch.distributed as dist

            if self.world > 1 and dist.is_available() and dist.is_initialized():
                t = torch.tensor([cap], dtype=torch.int64, device=self.device)
                dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)   # set-up time, once per buffer set
                cap = int(t.item())
        total = n_owned + 2 * (cap - 1)
        dev = self.device
        self._res = {
            "x": torch.empty(total, dtype=torch.float64, device=dev),
            "y": torch.empty(total, dtype=torch.float64, device=dev),
            "z": torch.empty(total, dtype=torch.float64, device=dev),
            "g": torch.empty(total, dtype=torch.int32, device=dev),
            "send": torch.zeros((2, cap, 4), dtype=torch.float64, device=dev),
            "recv": torch.zeros((2, cap, 4), dtype=torch.float64, device=dev),
            "counts": torch.zeros(2, dtype=torch.int32, device=dev),
            "total": torch.zeros(1, dtype=torch.int32, device=dev),
            "n_owned": int(n_owned), "cap": int(cap),
        }
        split = [0] * self.world
        split[self.left] += cap
        split[self.right] += cap
        self._res["split"] = split
        r = self._res
        return r["x"][:n_owned], r["y"][:n_owned], r["z"][:n_owned], r["g"][:n_owned]

    def resident_pack(self):
        """Boundary planes of the resident frame -> the two send buffers (device kernel, no host sync)."""
        torch = self.torch
        r = self._res
        n, cap = r["n_owned"], r["cap"]
        stream = torch.cuda.current_stream().cuda_stream
        b, o, p = L.box_args(self.box, self.origin, self.boundary)
        # all_to_all_single wants the segments in rank order: [left, right] or [right, left]
        li, ri = (0, 1) if self.left <= self.right else (1, 0)
        L.check(L.lib().mdb_slab_pack_device(
            C.c_void_p(r["x"].data_ptr()), C.c_void_p(r["y"].data_ptr()), C.c_void_p(r["z"].data_ptr()),
            C.c_void_p(r["g"].data_ptr()), n, L.dptr(b), L.dptr(o), L.iptr(p), self.rc, int(self.lo), int(self.hi),
            int(self.halo), C.c_void_p(r["send"][li].data_ptr()), C.c_void_p(r["send"][ri].data_ptr()), cap,
            C.c_void_p(r["counts"].data_ptr()), C.c_void_p(int(stream))))

    def resident_unpack(self, device_index: Optional[int] = None):
        """Received buffers -> ghosts behind the owned atoms; returns the loaded DeviceSystem."""
        torch = self.torch
        r = self._res
        n, cap = r["n_owned"], r["cap"]
        stream = torch.cuda.current_stream().cuda_stream
        L.check(L.lib().mdb_slab_unpack_device(
            C.c_void_p(r["recv"][0].data_ptr()), C.c_void_p(r["recv"][1].data_ptr()), cap,
            C.c_void_p(r["x"].data_ptr()), C.c_void_p(r["y"].data_ptr()), C.c_void_p(r["z"].data_ptr()),
            C.c_void_p(r["g"].data_ptr()), n, 2 * (cap - 1), C.c_void_p(r["total"].data_ptr()), C.c_void_p(int(stream))))
        total = int(r["total"].item())            # the one host read of the frame
        if total < 0:
            raise RuntimeError("halo buffers too small for this frame: call resident_buffers with a larger cap")
        self.n_owned = self.n_rows = n
        self.halo_atoms = total - n
        ds = self.device_system(device_index)
        ds.set_slab_device(r["x"][:total], r["y"][:total], r["z"][:total], r["g"][:total], n, self.plane0, self.nplanes,
                           self.box, self.origin, self.boundary, stream=stream)
        ds.set_local_fraction(min(1.0, self.nplanes / self.n0))
        self.local = (r["x"][:total], r["y"][:total], r["z"][:total], r["g"][:total])
        self.local_extra = None
        return ds

    def exchange_resident(self, device_index: Optional[int] = None):
        """Halo exchange of the resident frame; returns the DeviceSystem loaded with owned + ghost atoms."""
        import torch.distributed as dist

        r = self._res
        cap = r["cap"]
        self.resident_pack()
        dist.all_to_all_single(r["recv"].view(2 * cap, 4), r["send"].view(2 * cap, 4), output_split_sizes=r["split"],
                               input_split_sizes=r["split"], group=self.group)
        return self.resident_unpack(device_index)

    def make_step(self, x, y, z, gid, fused: bool = False):
        """Benchmark step: one frame = boundary pack + halo exchange + ghost append + binning + neighbour build
        + CNA (``fused``: the fused neighbour + CNA kernel, no list)."""
        if self.world == 1 or self.halo != 1:
            def step():
                ds = self.build(x, y, z, gid)
                ds.fcna(self.rc, fetch=False)
                return ds.M

            return step
        rx, ry, rz, rg = self.resident_buffers(int(x.shape[0]))
        rx.copy_(x)
        ry.copy_(y)
        rz.copy_(z)
        rg.copy_(gid.to(self.torch.int32))

        def step():
            ds = self.exchange_resident()
            if fused:
                lab, used = ds.fused_cna(self.rc, fetch=False)
                assert used
                return 0
            M, mx = ds.build_neighbor(self.rc, None)
            ds.fcna(self.rc, fetch=False)
            return M

        return step

    # ------------------------------------------------------------------ reductions
    def all_reduce_sum(self, arr: np.ndarray) -> np.ndarray:
        """Sum a small host array over the ranks (RDF pair counts)."""
        if self.world == 1:
            return arr
        import torch.distributed as dist

        t = self.torch.as_tensor(np.ascontiguousarray(arr), device=self.device)
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        return t.cpu().numpy()


# halo depth (in cut-off cell planes) a descriptor needs; see the module docstring
HALO_CUTOFF = {"neighbor": 1, "fcna": 1, "csp": 1, "aja": 1, "steinhardt": 1, "rdf": 1,
               "steinhardt_average": 2, "solid_liquid": 3, "solid_liquid_average": 4}
# levels of neighbour indirection of the k-nearest consumers (KnnDecomposition.verify)
KNN_LEVELS = {"csp": 1, "aja": 1, "acna": 1, "ptm": 1, "ptm_two_shell": 2, "ids": 4}


class KnnDecomposition(SlabDecomposition):
    """Slab decomposition for k-nearest lists.  The planes have the nominal width
    ``safety * (3 (k+1) / (4 pi rho))^(1/3)`` (rho = global number density), ``halo`` of them are
    exchanged per side, and :meth:`build_knn` verifies the result (module docstring)."""

    def __init__(self, box, origin, boundary, k: int, n_total: int, rank: int, world: int, device=None, group=None,
                 halo: int = 1, safety: float = 1.6, width: Optional[float] = None):
        box = np.asarray(box, float).reshape(3, 3)
        vol = abs(float(np.linalg.det(box)))
        if width is None:
            width = safety * (3.0 * (k + 1) / (4.0 * np.pi * (n_total / vol))) ** (1.0 / 3.0)
        self.k = int(k)
        self.n_total = int(n_total)
        super().__init__(box, origin, boundary, width, rank, world, device, group, halo)

    def plane_coordinate(self, x, y, z):
        """Continuous x plane coordinate tau (perpendicular distance from the box face / plane width) of
        wrapped positions; the integer plane is floor(tau) clamped to n0 - 1."""
        torch = self.torch
        o = torch.tensor(self.origin, dtype=torch.float64, device=x.device)
        hinv = torch.tensor(np.linalg.inv(self.box), dtype=torch.float64, device=x.device)
        # fractional coordinate along a: r . hinv[:, 0]
        f = (x - o[0]) * hinv[0, 0] + (y - o[1]) * hinv[1, 0] + (z - o[2]) * hinv[2, 0]
        if self.boundary[0]:
            f = f - torch.floor(f)
        thick = 1.0 / float(np.linalg.norm(np.linalg.inv(self.box)[:, 0]))
        return f * (thick / self.rc)

    def verify(self, ds, levels: int = 1, k: Optional[int] = None):
        """True iff every k-nearest row an owned result depends on (``levels`` hops of neighbour
        indirection) is complete: its k-th distance does not reach past the stored window."""
        torch = self.torch
        if self.world == 1 or self.nplanes >= self.n0:
            return True
        ax, ay, az, _ = self.local
        N = int(ax.shape[0])
        vptr, dptr, nptr, M = ds.neighbor_device()
        dist = torch.as_tensor(_DeviceView(dptr, (N, M), "<f8"), device=ax.device)
        dk = dist[:, int(k or M) - 1]
        short = dk < 0                                     # fewer than k atoms exist at all
        tau = self.plane_coordinate(ax, ay, az)
        u = tau - (self.lo - self.halo)                    # plane coordinate inside the stored window
        u = torch.where(u < 0, u + self.n0, u)
        u = torch.where(u >= self.n0, u - self.n0, u)
        inf = torch.full_like(u, float("inf"))
        periodic = bool(self.boundary[0])
        left = u * self.rc if (periodic or self.rank > 0) else inf
        right = (self.nplanes - u) * self.rc if (periodic or self.rank < self.world - 1) else inf
        edge = torch.minimum(left, right) * (1.0 - 1e-12)
        complete = (dk <= edge) & ~short
        # distance (lower bound) of every local atom from the owned plane range [halo, halo + owned)
        own_lo, own_hi = float(self.halo), float(self.halo + self.hi - self.lo)
        gap = torch.clamp(torch.maximum(own_lo - u, u - own_hi), min=0.0) * self.rc
        reach = 0.0
        for _ in range(int(levels)):
            need = gap <= reach
            if not bool(complete[need].all().item()):
                return False
            reach += float(dk[need].max().item())
        return True

    def build_knn(self, x, y, z, gid, k: Optional[int] = None, levels: int = 1, extra=None, max_halo: int = 8,
                  device_index: Optional[int] = None, replicated: bool = False, collective: bool = True):
        """k-nearest lists for the owned atoms.  All ranks widen the halo together until every rank's
        rows verify.  Returns the DeviceSystem (rows: owned + inner ghost layers)."""
        import torch.distributed as dist

        torch = self.torch
        k = int(k or self.k)
        assert k <= self.k, "the decomposition was sized for a smaller k"
        while True:
            ds = self.load(x, y, z, gid, extra, device_index, replicated)
            ds.build_knn(k)
            ok = self.verify(ds, levels, k)
            if self.world > 1 and collective:
                t = torch.tensor([1 if ok else 0], dtype=torch.int32, device=x.device)
                dist.all_reduce(t, op=dist.ReduceOp.MIN, group=self.group)
                ok = bool(t.item())
            if ok:
                return ds
            if self.halo >= max_halo or (2 * (self.halo + 1) + 1) * self.world > self.n0:
                raise RuntimeError(
                    f"k-nearest halo of {self.halo} planes (width {self.rc:.4g}) is not enough and cannot grow "
                    f"further on {self.world} ranks ({self.n0} planes)")
            self._widen()

    def _widen(self):
        self.halo += 1
        self.bounds = slab_bounds(self.n0, self.world, self.halo)
        self.plane0 = (self.lo - self.halo) % self.n0
        self.nplanes = self.hi - self.lo + 2 * self.halo
