#!/usr/bin/env python
"""Throughput of BASELINE configs[2] and configs[3] on N GPUs (run under torchrun, one rank per GPU; also
works as a plain 1-GPU process):

  C3  ~50 M-atom BCC Fe (a = 2.8665, n = 292, rattled sigma = 0.05): k-nearest(18) + PTM("fcc-hcp-bcc")
      + Ackland-Jones through KnnDecomposition (ghost-atom halo, verified)
  C4  ~20 M-atom thermal FCC Al stand-in for the 200-grain polycrystal (the polycrystal builder is a
      "next" row of SURVEY.md 8f): Steinhardt q4, q6 (rc = 0.85 a) + RDF 500 bins (rc = 6.0) with the pair
      counts all-reduced over the ranks

Each rank generates its own lattice planes (input distribution outside the timed region); one step =
halo exchange + binning + list build + descriptor, timed with CUDA events on the launching stream,
max over ranks, best of `--reps`.  Rank 0 prints one JSON line per kernel and writes them to
gpurun_out/config_throughput_<N>gpu.json (summarised in profiles/r1_config_throughput_multi.json)."""
from __future__ import annotations

import argparse
import json
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

BCC = [[0.0, 0.0, 0.0], [0.5, 0.5, 0.5]]
FCC = [[0.0, 0.0, 0.0], [0.5, 0.5, 0.0], [0.0, 0.5, 0.5], [0.5, 0.0, 0.5]]


def lattice_planes_dev(torch, basis, a, n, ix0, ix1, sigma, seed, dev):
    planes = sorted({i % n for i in range(ix0, ix1)})   # periodic: a rattled atom of plane 0 may wrap to the far side
    return _planes_dev(torch, basis, a, n, planes, sigma, seed, dev)


def _planes_dev(torch, basis, a, n, planes, sigma, seed, dev):
    """Lattice planes ix in [ix0, ix1) of an n^3 supercell, rattled with a per-plane seed (so every rank
    generates identical atoms for the planes it shares with others).  Returns x, y, z, global id."""
    b = torch.tensor(basis, dtype=torch.float64, device=dev) * a
    nb = b.shape[0]
    g = torch.arange(n, dtype=torch.float64, device=dev) * a
    xs, ys, zs, ids = [], [], [], []
    for ix in planes:
        x = (ix * a + b[:, 0].view(1, 1, nb)).expand(n, n, nb).reshape(-1)
        y = (g.view(n, 1, 1) + b[:, 1].view(1, 1, nb)).expand(n, n, nb).reshape(-1)
        z = (g.view(1, n, 1) + b[:, 2].view(1, 1, nb)).expand(n, n, nb).reshape(-1)
        if sigma > 0:
            gen = torch.Generator(device=dev).manual_seed(seed * 100003 + ix)
            r = torch.randn((3, x.numel()), dtype=torch.float64, device=dev, generator=gen) * sigma
            x, y, z = x + r[0], y + r[1], z + r[2]
        xs.append(x)
        ys.append(y)
        zs.append(z)
        ids.append(torch.arange(ix * n * n * nb, (ix + 1) * n * n * nb, dtype=torch.int32, device=dev))
    return torch.cat(xs), torch.cat(ys), torch.cat(zs), torch.cat(ids)


def main():
    import torch
    import torch.distributed as dist

    from mdapy_b200.distributed import KnnDecomposition, SlabDecomposition

    ap = argparse.ArgumentParser()
    ap.add_argument("--c3-cells", type=int, default=292)
    ap.add_argument("--c4-cells", type=int, default=171)
    ap.add_argument("--reps", type=int, default=2)
    ap.add_argument("--skip", default="")
    ap.add_argument("--c4-input", default="thermal", choices=["thermal", "polycrystal"],
                    help="polycrystal: 200-grain periodic Voronoi polycrystal (tools/polycrystal.py), L = 692.5 A")
    ap.add_argument("--poly-L", type=float, default=692.5)
    ap.add_argument("--poly-grains", type=int, default=200)
    args = ap.parse_args()
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    o, bnd = np.zeros(3), [1, 1, 1]
    rows = []

    def timed(fn):
        best = None
        for _ in range(args.reps):
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            if world > 1:
                t = torch.tensor([ms], dtype=torch.float64, device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                ms = float(t.item())
            best = ms if best is None else min(best, ms)
        return best

    def report(config, kernel, n_atoms, ms, **kw):
        r = {"config": config, "kernel": kernel, "n_gpus": world, "atoms": n_atoms, "ms": ms,
             "atoms_per_s": n_atoms / (ms * 1e-3), **kw}
        rows.append(r)
        if rank == 0:
            print(json.dumps(r), flush=True)

    def own_planes(dec, n, a):
        """Lattice planes covering this rank's slab, then keep exactly the owned atoms."""
        ix0, ix1 = dec.lattice_planes(n, a) if world > 1 else (0, n)
        return ix0, ix1

    def keep_owned(dec, x, y, z, gid):
        if world == 1:
            return x, y, z, gid
        pl = dec.planes(x, y, z)
        k = (pl >= dec.lo) & (pl < dec.hi)
        return x[k].contiguous(), y[k].contiguous(), z[k].contiguous(), gid[k].contiguous()

    # ------------------------------------------------------------------ C3: BCC Fe, kNN(18) + PTM + AJA
    if "c3" not in args.skip:
        a, n = 2.8665, args.c3_cells
        box = np.diag([n * a] * 3).astype(float)
        N = 2 * n ** 3
        dec = KnnDecomposition(box, o, bnd, 18, N, rank, world, dev)
        ix0, ix1 = own_planes(dec, n, a)
        # rattled atoms may cross a plane boundary: generate one extra lattice plane on each side
        x, y, z, gid = lattice_planes_dev(torch, BCC, a, n, ix0 - 1, ix1 + 1, 0.05, 2, dev)
        x, y, z, gid = keep_owned(dec, x, y, z, gid)
        cnt = torch.tensor([x.numel()], dtype=torch.int64, device=dev)
        if world > 1:
            dist.all_reduce(cnt)
        assert int(cnt.item()) == N, (int(cnt.item()), N)
        state = {}

        def knn18():
            state["ds"] = dec.build_knn(x, y, z, gid, 18)

        t_knn = timed(knn18)
        ds = state["ds"]
        t_ptm = timed(lambda: ds.ptm("fcc-hcp-bcc", 0.1, fetch=False))
        out, _ = ds.ptm("fcc-hcp-bcc", 0.1)
        frac_bcc = float((out[: dec.n_owned, 0] == 3).mean())

        def knn14():
            state["ds"] = dec.build_knn(x, y, z, gid, 14)

        t_knn14 = timed(knn14)
        ds = state["ds"]
        t_aja = timed(lambda: ds.aja(fetch=False))
        report("C3 BCC Fe", "halo exchange + k-nearest(18)", N, t_knn, knn_halo=dec.halo, halo_atoms=dec.halo_atoms)
        report("C3 BCC Fe", "PTM fcc-hcp-bcc", N, t_ptm, bcc_fraction_rank0=frac_bcc)
        report("C3 BCC Fe", "halo exchange + k-nearest(14)", N, t_knn14)
        report("C3 BCC Fe", "Ackland-Jones", N, t_aja)
        report("C3 BCC Fe", "pipeline kNN18 + PTM + kNN14 + AJA", N, t_knn + t_ptm + t_knn14 + t_aja)
        del ds, state, x, y, z, gid, dec
        torch.cuda.empty_cache()

    # ------------------------------------------------------------------ C4: Steinhardt q4/q6 + RDF 500 bins
    if "c4" not in args.skip:
        a, n = 4.05, args.c4_cells
        poly = args.c4_input == "polycrystal"
        cfg = "C4 FCC Al thermal"
        if poly:
            # every rank generates the same frame (deterministic) and selects its slab locally (replicated input)
            sys.path.insert(0, str(ROOT / "tools"))
            from polycrystal import make_polycrystal

            Lp = args.poly_L
            box = np.diag([Lp] * 3).astype(float)
            px, py, pz, _ = make_polycrystal(Lp, args.poly_grains, a, 7, dev)
            pgid = torch.arange(px.numel(), dtype=torch.int32, device=dev)
            N = int(px.numel())
            cfg = f"C4 {args.poly_grains}-grain FCC Al polycrystal L={Lp}"
        else:
            box = np.diag([n * a] * 3).astype(float)
            N = 4 * n ** 3
        rc_q, rc_g, nbin = 0.85 * a, 6.0, 500
        for name, rc, halo in (("steinhardt", rc_q, 1), ("steinhardt_average", rc_q, 2), ("rdf", rc_g, 1)):
            dec = SlabDecomposition(box, o, bnd, rc, rank, world, dev, halo=halo)
            if poly:
                x, y, z, gid = px, py, pz, pgid
            else:
                ix0, ix1 = own_planes(dec, n, a)
                x, y, z, gid = lattice_planes_dev(torch, FCC, a, n, ix0 - 1, ix1 + 1, 0.12, 4, dev)
                x, y, z, gid = keep_owned(dec, x, y, z, gid)
            state = {}

            def build():
                state["ds"] = dec.build(x, y, z, gid, sync_width=True, replicated=poly)

            t_b = timed(build)
            ds = state["ds"]
            if name == "rdf":
                types = np.zeros(ds.N, np.int32)
                res = {}

                def rdf():
                    res["g"] = dec.all_reduce_sum(ds.rdf_counts(rc, nbin, type_list=types, ntype=1))

                t_k = timed(rdf)
                report(cfg, f"neighbour build rc={rc}", N, t_b, M=ds.M)
                report(cfg, "RDF 500 bins from the list + all-reduce", N, t_k, pairs=float(res["g"].sum()))
            else:
                avg = name.endswith("average")
                t_k = timed(lambda: ds.steinhardt([4, 6], rc=rc, average=avg, fetch=False))
                extra = {}
                if poly and not avg:
                    q, _, _ = ds.steinhardt([4, 6], rc=rc, average=False)
                    q6 = q[: dec.n_owned, 1]
                    extra = {"q6_mean_rank0": float(np.nanmean(q6)), "fcc_like_fraction_rank0": float((q6 > 0.5).mean())}
                report(cfg, f"neighbour build rc={rc:.4f} halo={halo}", N, t_b, M=ds.M)
                report(cfg, "Steinhardt q4,q6" + (" averaged" if avg else ""), N, t_k, **extra)
            del ds, state, dec
            if not poly:
                del x, y, z, gid
            torch.cuda.empty_cache()

    if rank == 0:
        (ROOT / "gpurun_out").mkdir(exist_ok=True)   # merged back by gpurun; copied into profiles/ afterwards
        (ROOT / "gpurun_out" / f"config_throughput_{world}gpu.json").write_text(json.dumps(rows, indent=1))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
