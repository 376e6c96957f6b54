"""CPU test of the multi-rank host logic (world_size 2, gloo): slab bounds, migration to owners,
ghost-plane exchange.  Plane ids come from NumPy here (ortho cell, origin 0); on the GPU they come
from the library kernel with the reference's cell arithmetic (tests/test_gpu_slab.py)."""
import os
import socket

import numpy as np
import pytest

import helpers as H


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, pos, box, rc, out, halo=1):
    import torch
    import torch.distributed as dist

    from mdapy_b200.distributed import SlabDecomposition

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        dec = SlabDecomposition(box, np.zeros(3), [1, 1, 1], rc, rank, world, halo=halo)
        n0 = dec.n0
        # every rank starts with an arbitrary interleaved share of the atoms
        mine = np.arange(rank, pos.shape[0], world)
        x, y, z = (torch.tensor(pos[mine, k]) for k in range(3))
        gid = torch.tensor(mine.astype(np.int32))

        def planes_of(xx):
            return torch.clamp(torch.floor(xx * (1.0 / rc)), 0, n0 - 1).to(torch.int32)

        x, y, z, gid = dec.migrate(x, y, z, gid, planes=planes_of(x))
        pl = planes_of(x)
        assert bool(((pl >= dec.lo) & (pl < dec.hi)).all())
        types = (gid % 3).to(torch.int32)
        gx, gy, gz, gg, (gt,) = dec.exchange_halo(x, y, z, gid, pl, extra=[types])
        assert torch.equal(gt, (gg % 3).to(torch.int32)), "per-atom payload must travel with its atom"
        gpl = planes_of(gx)
        want = {(dec.lo - 1 - h) % n0 for h in range(halo)} | {(dec.hi + h) % n0 for h in range(halo)}
        assert set(np.unique(gpl.numpy()).tolist()) == want
        layer = dec.ghost_layer(gpl)
        assert int(layer.min()) >= 1 and int(layer.max()) == halo
        out[rank] = (gid.numpy().copy(), gg.numpy().copy(), gx.numpy().copy())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("halo", [1, 2])
def test_migrate_and_halo_two_ranks(halo):
    import torch.multiprocessing as mp

    pos, box = H.fcc(3.615, 10 if halo == 2 else 8)
    pos = H.rattle(pos, 0.05, 0) % np.diag(box)
    rc = 3.615 * 0.8536
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), pos, box, rc, out, halo), nprocs=world, join=True)
    owned = np.concatenate([out[r][0] for r in range(world)])
    assert sorted(owned.tolist()) == list(range(pos.shape[0])), "every atom must be owned exactly once"
    n0 = int(np.floor(box[0, 0] / rc))
    planes = np.clip(np.floor(pos[:, 0] / rc), 0, n0 - 1).astype(int)
    for r in range(world):
        lo, hi = r * n0 // world, (r + 1) * n0 // world
        gp = [(lo - 1 - h) % n0 for h in range(halo)] + [(hi + h) % n0 for h in range(halo)]
        ghosts_expected = np.nonzero(np.isin(planes, gp))[0]
        assert sorted(out[r][1].tolist()) == sorted(ghosts_expected.tolist())
        # raw coordinates travel unchanged
        assert np.array_equal(np.sort(out[r][2]), np.sort(pos[ghosts_expected, 0]))


def test_slab_bounds_reject_small_grids():
    from mdapy_b200.distributed import slab_bounds

    assert slab_bounds(12, 4) == [0, 3, 6, 9, 12]
    with pytest.raises(ValueError):
        slab_bounds(5, 2)
    assert slab_bounds(10, 2, halo=2) == [0, 5, 10]
    with pytest.raises(ValueError):
        slab_bounds(9, 2, halo=2)
