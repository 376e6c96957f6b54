"""GPU parity of the decomposed descriptors: every emulated rank selects its slab + halo
from the replicated frame (no process group needed), runs the ordinary kernels, and its OWNED rows must
equal the single-GPU result -- bit for bit for list-order-dependent sums (Steinhardt), exactly for labels
and pair counts.  The NCCL exchange itself is covered on CPU by tests/test_distributed_gloo.py and on
GPUs by tools/dist_check.py (torchrun)."""
import numpy as np
import pytest

import helpers as H

pytestmark = pytest.mark.gpu


def _torch_frame(pos):
    import torch

    dev = torch.device("cuda", 0)
    x, y, z = (torch.tensor(np.ascontiguousarray(pos[:, k]), device=dev) for k in range(3))
    gid = torch.arange(pos.shape[0], dtype=torch.int32, device=dev)
    return dev, x, y, z, gid


def _single(pos, box, boundary):
    from mdapy_b200.device import DeviceSystem

    ds = DeviceSystem(0)
    ds.set_atoms(pos[:, 0], pos[:, 1], pos[:, 2], box, np.zeros(3), boundary)
    return ds


@pytest.mark.parametrize("world,halo,average", [(2, 2, True), (3, 2, True), (2, 1, False)])
def test_steinhardt_on_slabs(world, halo, average):
    from mdapy_b200.distributed import SlabDecomposition

    p, b = H.fcc(4.05, 16)
    pos = H.rattle(p, 0.08, 11)
    rc = 0.85 * 4.05
    bnd = [1, 1, 1]
    full = _single(pos, b, bnd)
    full.build_neighbor(rc)
    ref, _, _ = full.steinhardt([4, 6], rc=rc, average=average, wl=True)
    dev, x, y, z, gid = _torch_frame(pos)
    seen = 0
    for r in range(world):
        dec = SlabDecomposition(b, np.zeros(3), bnd, rc, r, world, dev, halo=halo)
        ds = dec.build(x, y, z, gid, replicated=True)
        assert dec.n_rows >= dec.n_owned and ds.n_rows == dec.n_rows
        q, _, _ = ds.steinhardt([4, 6], rc=rc, average=average, wl=True)
        rows = dec.local[3][: dec.n_owned].cpu().numpy()
        assert np.array_equal(q[: dec.n_owned].view(np.int64), ref[rows].view(np.int64)), f"rank {r}"
        seen += rows.size
    assert seen == pos.shape[0]


@pytest.mark.parametrize("average,halo", [(False, 3), (True, 4)])
def test_solid_liquid_on_slabs(average, halo):
    from mdapy_b200.distributed import SlabDecomposition

    p, b = H.fcc(4.05, 24)
    rng = np.random.default_rng(3)
    pos = p.copy()
    half = pos[:, 1] > 0.5 * b[1, 1]          # one half strongly disordered: a solid / liquid-like interface
    pos[half] += rng.normal(0, 0.45, (int(half.sum()), 3))
    pos[~half] += rng.normal(0, 0.05, (int((~half).sum()), 3))
    rc = 0.85 * 4.05
    bnd = [1, 1, 1]
    full = _single(pos, b, bnd)
    full.build_neighbor(rc)
    full.steinhardt([6], rc=rc, average=average)
    ref_sl, ref_nb = full.solid_liquid(0, 0.7, 7, rc=rc)
    assert 0 < ref_sl.sum() < pos.shape[0]
    dev, x, y, z, gid = _torch_frame(pos)
    world = 2
    for r in range(world):
        dec = SlabDecomposition(b, np.zeros(3), bnd, rc, r, world, dev, halo=halo)
        ds = dec.build(x, y, z, gid, replicated=True)
        ds.steinhardt([6], rc=rc, average=average)
        sl, nb = ds.solid_liquid(0, 0.7, 7, rc=rc)
        rows = dec.local[3][: dec.n_owned].cpu().numpy()
        assert np.array_equal(nb[: dec.n_owned], ref_nb[rows])
        assert np.array_equal(sl[: dec.n_owned], ref_sl[rows])


@pytest.mark.parametrize("world", [2, 3])
def test_rdf_counts_sum_over_slabs(world):
    import torch

    from mdapy_b200.distributed import SlabDecomposition

    p, b = H.fcc(4.05, 14)
    pos = H.rattle(p, 0.1, 5)
    types = (np.random.default_rng(1).random(pos.shape[0]) < 0.3).astype(np.int32)
    rc, nbin = 3.9, 60
    bnd = [1, 1, 1]
    full = _single(pos, b, bnd)
    full.build_neighbor(rc)
    ref_typed = full.rdf_counts(rc, nbin, type_list=types, ntype=2)
    ref_single = full.rdf_counts(rc, nbin)
    ref_stream = full.rdf_counts(rc, nbin, type_list=types, ntype=2, streaming=True)
    assert np.array_equal(ref_typed, ref_stream)
    dev, x, y, z, gid = _torch_frame(pos)
    tt = torch.tensor(types, device=dev)
    acc_t, acc_s, acc_st = np.zeros_like(ref_typed), np.zeros_like(ref_single), np.zeros_like(ref_typed)
    for r in range(world):
        dec = SlabDecomposition(b, np.zeros(3), bnd, rc, r, world, dev)
        ds = dec.build(x, y, z, gid, replicated=True, extra=[tt])
        lt = dec.local_extra[0].cpu().numpy()
        acc_t += ds.rdf_counts(rc, nbin, type_list=lt, ntype=2)
        acc_s += ds.rdf_counts(rc, nbin)
        acc_st += ds.rdf_counts(rc, nbin, type_list=lt, ntype=2, streaming=True)
    assert np.array_equal(acc_t, ref_typed)
    assert np.array_equal(acc_s, ref_single)
    assert np.array_equal(acc_st, ref_typed)


def _knn_frame(kind):
    if kind == "bcc":
        p, b = H.bcc(2.8665, 22)
        return H.rattle(p, 0.06, 2), b
    p, b = H.fcc(3.615, 18)
    return H.rattle(p, 0.07, 4), b


@pytest.mark.parametrize("kind,world", [("bcc", 2), ("fcc", 3)])
def test_knn_descriptors_on_slabs(kind, world):
    from mdapy_b200.distributed import KnnDecomposition

    pos, b = _knn_frame(kind)
    bnd = [1, 1, 1]
    N = pos.shape[0]
    full = _single(pos, b, bnd)
    full.build_knn(18)
    fv, fd, _ = full.fetch_neighbor()
    ref_ptm, _ = full.ptm("fcc-hcp-bcc", 0.1)
    full.build_knn(14)
    ref_aja = full.aja()
    ref_acna = full.acna()
    full.build_knn(12)
    ref_csp = full.csp(12)
    dev, x, y, z, gid = _torch_frame(pos)
    seen = 0
    for r in range(world):
        dec = KnnDecomposition(b, np.zeros(3), bnd, 18, N, r, world, dev)
        ds = dec.build_knn(x, y, z, gid, 18, replicated=True, collective=False)
        rows = dec.local[3][: dec.n_owned].cpu().numpy()
        v, d, _ = ds.fetch_neighbor()
        assert np.array_equal(d[: dec.n_owned].view(np.int64), fd[rows].view(np.int64)), "k-nearest distances differ"
        assert np.array_equal(v[: dec.n_owned], fv[rows]), "k-nearest indices (global ids) differ"
        out, _ = ds.ptm("fcc-hcp-bcc", 0.1)
        assert np.array_equal(out[: dec.n_owned, 0], ref_ptm[rows, 0])
        assert np.allclose(out[: dec.n_owned, 2], ref_ptm[rows, 2], rtol=0, atol=1e-12)
        ds = dec.build_knn(x, y, z, gid, 14, replicated=True, collective=False)
        assert np.array_equal(ds.aja()[: dec.n_owned], ref_aja[rows])
        assert np.array_equal(ds.acna()[: dec.n_owned], ref_acna[rows])
        ds = dec.build_knn(x, y, z, gid, 12, replicated=True, collective=False)
        got = ds.csp(12)[: dec.n_owned]
        assert np.array_equal(got.view(np.int64), ref_csp[rows].view(np.int64))
        seen += rows.size
    assert seen == N


def test_knn_halo_widens_when_too_thin():
    """A deliberately thin nominal plane width must be detected by verify() and widened."""
    from mdapy_b200.distributed import KnnDecomposition

    pos, b = _knn_frame("fcc")
    bnd = [1, 1, 1]
    full = _single(pos, b, bnd)
    full.build_knn(12)
    fv, fd, _ = full.fetch_neighbor()
    dev, x, y, z, gid = _torch_frame(pos)
    dec = KnnDecomposition(b, np.zeros(3), bnd, 12, pos.shape[0], 0, 2, dev, width=1.3)
    assert dec.halo == 1
    ds = dec.build_knn(x, y, z, gid, 12, replicated=True, collective=False)
    assert dec.halo > 1, "a 1.3 A halo cannot hold 12 neighbours: the halo must have been widened"
    rows = dec.local[3][: dec.n_owned].cpu().numpy()
    v, d, _ = ds.fetch_neighbor()
    assert np.array_equal(d[: dec.n_owned].view(np.int64), fd[rows].view(np.int64))


def test_diamond_identification_on_slabs():
    from mdapy_b200.distributed import KNN_LEVELS, KnnDecomposition

    a = 5.43
    n = 16
    basis = np.array([[0, 0, 0], [0, .5, .5], [.5, 0, .5], [.5, .5, 0], [.25, .25, .25], [.25, .75, .75],
                      [.75, .25, .75], [.75, .75, .25]])
    cells = np.stack(np.meshgrid(np.arange(n), np.arange(n), np.arange(n), indexing="ij"), -1).reshape(-1, 3)
    pos = ((cells[:, None, :] + basis[None, :, :]) * a).reshape(-1, 3)
    b = np.diag([n * a] * 3)
    rng = np.random.default_rng(8)
    pos = pos + rng.normal(0, 0.05, pos.shape)
    blob = np.linalg.norm(pos - 0.5 * n * a, axis=1) < 9.0
    pos[blob] += rng.normal(0, 0.5, (int(blob.sum()), 3))      # a disordered inclusion: all seven labels appear
    bnd = [1, 1, 1]
    full = _single(pos, b, bnd)
    full.build_knn(4)
    ref = full.ids()
    assert len(np.unique(ref)) >= 3
    dev, x, y, z, gid = _torch_frame(pos)
    for r in range(2):
        dec = KnnDecomposition(b, np.zeros(3), bnd, 4, pos.shape[0], r, 2, dev, halo=3)
        ds = dec.build_knn(x, y, z, gid, 4, levels=KNN_LEVELS["ids"], replicated=True, collective=False)
        rows = dec.local[3][: dec.n_owned].cpu().numpy()
        assert np.array_equal(ds.ids()[: dec.n_owned], ref[rows])
