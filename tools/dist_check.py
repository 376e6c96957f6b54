#!/usr/bin/env python
"""Multi-GPU parity check of the decomposed path (run under torchrun, one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tools/dist_check.py [--cells 40]

Every rank starts with an arbitrary interleaved share of a rattled crystal, the frame is migrated to its
slab owners over NCCL, halos are exchanged, and every descriptor is compared -- on each rank, for its OWNED
atoms -- with the single-GPU result that the rank computes itself from the full frame:
cut-off list (rows, order, distances, global width), CNA, CSP, Ackland-Jones, Steinhardt (+ averaging),
identifySolidLiquid, RDF (all-reduced pair counts), k-nearest lists + PTM + adaptive CNA.
Prints one JSON line per rank-0 with the verdicts and exits non-zero on any mismatch."""
from __future__ import annotations

import argparse
import json
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))


def main():
    import torch
    import torch.distributed as dist

    import helpers as H
    from mdapy_b200.device import DeviceSystem
    from mdapy_b200.distributed import KnnDecomposition, SlabDecomposition

    ap = argparse.ArgumentParser()
    ap.add_argument("--cells", dest="n", type=int, default=40, help="FCC supercell edge (40 -> 256,000 atoms)")
    args = ap.parse_args()
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)

    a = 4.05
    p, box = H.fcc(a, args.n)
    pos = H.rattle(p, 0.08, 7)
    N = pos.shape[0]
    rng = np.random.default_rng(2)
    hot = pos[:, 1] > 0.6 * box[1, 1]
    pos[hot] += rng.normal(0, 0.4, (int(hot.sum()), 3))
    types = (rng.random(N) < 0.25).astype(np.int32)
    rc = 0.85 * a
    o, bnd = np.zeros(3), [1, 1, 1]
    verdict = {}

    # ---- single-GPU reference on every rank
    full = DeviceSystem(local)
    full.set_atoms(pos[:, 0], pos[:, 1], pos[:, 2], box, o, bnd)
    M_ref, _ = full.build_neighbor(rc)
    fv, fd, fn = full.fetch_neighbor()
    ref_cna = full.fcna(rc)
    ref_q, _, _ = full.steinhardt([4, 6], rc=rc, average=False)
    ref_rdf = full.rdf_counts(rc, 50, type_list=types, ntype=2)
    ref_qa, _, _ = full.steinhardt([4, 6], rc=rc, average=True)
    full.steinhardt([6], rc=rc, average=False)
    ref_sl, ref_nb = full.solid_liquid(0, 0.7, 7, rc=rc)
    full.build_knn(18)
    kv, kd, _ = full.fetch_neighbor()
    ref_ptm, _ = full.ptm("fcc-hcp-bcc", 0.1)
    full.build_knn(14)
    ref_acna, ref_aja = full.acna(), full.aja()
    full.build_knn(12)
    ref_csp = full.csp(12)
    del full

    # ---- the distributed path: interleaved input -> migrate -> halo -> kernels
    mine = np.arange(rank, N, world)
    x, y, z = (torch.tensor(np.ascontiguousarray(pos[mine, k]), device=dev) for k in range(3))
    gid = torch.tensor(mine.astype(np.int32), device=dev)
    tt = torch.tensor(types[mine], device=dev)

    def check(name, ok):
        verdict[name] = bool(ok)

    for halo, tag in ((1, "halo1"), (2, "halo2"), (3, "halo3")):
        dec = SlabDecomposition(box, o, bnd, rc, rank, world, dev, halo=halo)
        mx, my, mz, mg, (mt,) = dec.migrate(x, y, z, gid, extra=[tt])
        ds = dec.build(mx, my, mz, mg, extra=[mt], sync_width=True)
        rows = mg.cpu().numpy()
        n_own = dec.n_owned
        if halo == 1:
            v, d, n = ds.fetch_neighbor()
            check("list_width_global", ds.M == M_ref)
            check("list_rows", np.array_equal(v[:n_own], fv[rows]) and np.array_equal(n[:n_own], fn[rows]) and
                  np.array_equal(d[:n_own].view(np.int64), fd[rows].view(np.int64)))
            check("fcna", np.array_equal(ds.fcna(rc)[:n_own], ref_cna[rows]))
            q, _, _ = ds.steinhardt([4, 6], rc=rc, average=False)
            check("steinhardt", np.array_equal(q[:n_own].view(np.int64), ref_q[rows].view(np.int64)))
            lt = dec.local_extra[0].cpu().numpy()
            counts = dec.all_reduce_sum(ds.rdf_counts(rc, 50, type_list=lt, ntype=2))
            check("rdf_allreduce", np.array_equal(counts, ref_rdf))
            # resident fast path: device-side boundary pack, fixed-capacity exchange, ghost append (no host
            # round trip besides the atom count), list + CNA and the fused kernel on the slab
            if world > 1:
                rdec = SlabDecomposition(box, o, bnd, rc, rank, world, dev, halo=1)
                rx, ry, rz, rg = rdec.resident_buffers(int(mx.shape[0]))
                rx.copy_(mx), ry.copy_(my), rz.copy_(mz), rg.copy_(mg)
                for rep in range(2):   # the buffers are reused frame after frame
                    rs = rdec.exchange_resident()
                    rs.build_neighbor(rc, M_ref)
                    v2, d2, n2 = rs.fetch_neighbor()
                    check(f"resident_rows_{rep}", np.array_equal(v2[:n_own], fv[rows]) and np.array_equal(n2[:n_own], fn[rows])
                          and np.array_equal(d2[:n_own].view(np.int64), fd[rows].view(np.int64)))
                    check(f"resident_fcna_{rep}", np.array_equal(rs.fcna(rc)[:n_own], ref_cna[rows]))
                    lab, used = rs.fused_cna(rc)
                    check(f"resident_fused_cna_{rep}", used and np.array_equal(lab[:n_own], ref_cna[rows]))
        elif halo == 2:
            q, _, _ = ds.steinhardt([4, 6], rc=rc, average=True)
            check("steinhardt_average", np.array_equal(q[:n_own].view(np.int64), ref_qa[rows].view(np.int64)))
        else:
            ds.steinhardt([6], rc=rc, average=False)
            sl, nb = ds.solid_liquid(0, 0.7, 7, rc=rc)
            check("solid_liquid", np.array_equal(sl[:n_own], ref_sl[rows]) and np.array_equal(nb[:n_own], ref_nb[rows]))

    kdec = KnnDecomposition(box, o, bnd, 18, N, rank, world, dev)
    mx, my, mz, mg = kdec.migrate(x, y, z, gid)
    rows = mg.cpu().numpy()
    ds = kdec.build_knn(mx, my, mz, mg, 18)
    n_own = kdec.n_owned
    v, d, _ = ds.fetch_neighbor()
    check("knn18", np.array_equal(d[:n_own].view(np.int64), kd[rows].view(np.int64)) and np.array_equal(v[:n_own], kv[rows]))
    out, _ = ds.ptm("fcc-hcp-bcc", 0.1)
    check("ptm", np.array_equal(out[:n_own, 0], ref_ptm[rows, 0]) and
          np.allclose(out[:n_own, 2], ref_ptm[rows, 2], rtol=0, atol=1e-12))
    ds = kdec.build_knn(mx, my, mz, mg, 14)
    check("acna", np.array_equal(ds.acna()[:n_own], ref_acna[rows]))
    check("aja", np.array_equal(ds.aja()[:n_own], ref_aja[rows]))
    ds = kdec.build_knn(mx, my, mz, mg, 12)
    check("csp", np.array_equal(ds.csp(12)[:n_own].view(np.int64), ref_csp[rows].view(np.int64)))

    ok = torch.tensor([1 if all(verdict.values()) else 0], dtype=torch.int32, device=dev)
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    gathered = [None] * world
    dist.all_gather_object(gathered, verdict)
    if rank == 0:
        merged = {k: all(g[k] for g in gathered) for k in verdict}
        print(json.dumps({"dist_check": "ok" if bool(ok.item()) else "MISMATCH", "world": world, "atoms": N,
                          "knn_halo": kdec.halo, "checks": merged}), flush=True)
    dist.destroy_process_group()
    sys.exit(0 if bool(ok.item()) else 1)


if __name__ == "__main__":
    main()
