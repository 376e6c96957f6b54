"""GPU parity of the decomposed (slab) build: each emulated rank's rows must equal the corresponding
rows of the single-GPU build bit for bit -- membership, ORDER, distances, counts, CNA labels.
(The NCCL exchange itself is covered by tests/test_distributed_gloo.py on CPU and by bench.py --gpus N.)"""
import numpy as np
import pytest

import helpers as H

pytestmark = pytest.mark.gpu


def _full_and_slabs(pos, box, boundary, rc, world):
    import torch

    from mdapy_b200.device import DeviceSystem
    from mdapy_b200.distributed import SlabDecomposition

    dev = torch.device("cuda", 0)
    o = np.zeros(3)
    x, y, z = (torch.tensor(np.ascontiguousarray(pos[:, k]), device=dev) for k in range(3))
    full = DeviceSystem(0)
    full.set_atoms(pos[:, 0], pos[:, 1], pos[:, 2], box, o, boundary)
    full.build_neighbor(rc)
    fv, fd, fn = full.fetch_neighbor()
    fcna = full.fcna(rc)
    gid = torch.arange(pos.shape[0], dtype=torch.int32, device=dev)
    seen = 0
    for r in range(world):
        dec = SlabDecomposition(box, o, boundary, rc, r, world, dev)
        pl = dec.planes(x, y, z)
        own = (pl >= dec.lo) & (pl < dec.hi)
        ghost = (pl == (dec.lo - 1) % dec.n0) | (pl == dec.hi % dec.n0)
        # shuffle ghosts and owned independently: local order must not matter
        oi = torch.nonzero(own).flatten()
        gi = torch.nonzero(ghost & ~own).flatten()
        oi = oi[torch.randperm(oi.numel(), device=dev, generator=torch.Generator(device=dev).manual_seed(r))]
        gi = gi[torch.randperm(gi.numel(), device=dev, generator=torch.Generator(device=dev).manual_seed(r + 9))]
        sel = torch.cat([oi, gi])
        ds = DeviceSystem(0)
        ds.set_slab_device(x[sel].contiguous(), y[sel].contiguous(), z[sel].contiguous(), gid[sel].contiguous(),
                           int(oi.numel()), dec.plane0, dec.nplanes, box, o, boundary)
        M, mx = ds.build_neighbor(rc)
        v, d, n = ds.fetch_neighbor()
        rows = oi.cpu().numpy()
        k = min(M, fv.shape[1])
        assert np.array_equal(n, fn[rows])
        assert mx <= fv.shape[1]
        assert np.array_equal(v[:, :k], fv[rows][:, :k]), f"rank {r}: rows differ from the single-GPU build"
        assert np.array_equal(d[:, :k].view(np.int64), fd[rows][:, :k].view(np.int64))
        assert np.array_equal(ds.fcna(rc), fcna[rows])
        lab, used = ds.fused_cna(rc)          # fused kernel on the slab window (eligible frames only)
        if used:
            assert np.array_equal(lab, fcna[rows]), f"rank {r}: fused labels differ"
        seen += rows.size
    assert seen == pos.shape[0]


@pytest.mark.parametrize("world", [2, 3])
def test_slab_rows_equal_single_gpu(world):
    p, b = H.fcc(3.615, 12)
    _full_and_slabs(H.rattle(p, 0.07, world), b, [1, 1, 1], 3.615 * 0.8536, world)


def test_slab_triclinic_mixed_boundary():
    p, b = H.fcc(3.615, 12)
    ps, bs = H.shear(H.rattle(p, 0.05, 5), b, xy=0.15, xz=0.05, yz=-0.1)
    _full_and_slabs(ps, bs, [1, 1, 0], 3.2, 2)


def test_slab_open_x_axis():
    p, b = H.fcc(3.615, 12)
    _full_and_slabs(H.rattle(p, 0.05, 6), b, [0, 1, 1], 3.3, 2)


@pytest.mark.parametrize("world", [2, 4])
def test_resident_pack_unpack_emulated_ranks(world):
    """Device-side halo exchange (boundary pack kernel, fixed-capacity buffers, ghost append): the NCCL
    all-to-all is replaced by copying the send buffers between emulated ranks on one GPU."""
    import torch

    from mdapy_b200.device import DeviceSystem
    from mdapy_b200.distributed import SlabDecomposition

    p, b = H.fcc(3.615, 16)
    pos = H.rattle(p, 0.07, 3)
    rc, o, bnd = 3.615 * 0.8536, np.zeros(3), [1, 1, 1]
    dev = torch.device("cuda", 0)
    x, y, z = (torch.tensor(np.ascontiguousarray(pos[:, k]), device=dev) for k in range(3))
    gid = torch.arange(pos.shape[0], dtype=torch.int32, device=dev)
    full = DeviceSystem(0)
    full.set_atoms(pos[:, 0], pos[:, 1], pos[:, 2], b, o, bnd)
    M, _ = full.build_neighbor(rc)
    fv, fd, fn = full.fetch_neighbor()
    fcna = full.fcna(rc)
    decs, owned = [], []
    for r in range(world):
        dec = SlabDecomposition(b, o, bnd, rc, r, world, dev)
        pl = dec.planes(x, y, z)
        oi = torch.nonzero((pl >= dec.lo) & (pl < dec.hi)).flatten()
        rx, ry, rz, rg = dec.resident_buffers(int(oi.numel()), cap=3 * pos.shape[0] // dec.n0)   # one cap for all ranks
        rx.copy_(x[oi]), ry.copy_(y[oi]), rz.copy_(z[oi]), rg.copy_(gid[oi])
        decs.append(dec)
        owned.append(oi.cpu().numpy())
    for frame in range(2):     # buffers are reused frame after frame
        for dec in decs:
            dec.resident_pack()
        for r, dec in enumerate(decs):     # what the all-to-all delivers: the left neighbour's RIGHT buffer, ...
            left, right = decs[dec.left], decs[dec.right]

            def buf(d, which):   # which: "left" / "right" send buffer of rank d
                li, ri = (0, 1) if d.left <= d.right else (1, 0)
                return d._res["send"][li if which == "left" else ri]

            dec._res["recv"][0].copy_(buf(left, "right"))
            dec._res["recv"][1].copy_(buf(right, "left"))
        for r, dec in enumerate(decs):
            ds = dec.resident_unpack(0)
            ds.build_neighbor(rc, M)
            v, d, n = ds.fetch_neighbor()
            rows = owned[r]
            assert np.array_equal(n, fn[rows]) and np.array_equal(v, fv[rows])
            assert np.array_equal(d.view(np.int64), fd[rows].view(np.int64))
            lab, used = ds.fused_cna(rc)
            assert used and np.array_equal(lab, fcna[rows])
