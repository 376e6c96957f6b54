"""oracle/pipeline.py -- TEST INFRASTRUCTURE ONLY (never imported by mdapy_b200).

NumPy restatement of mdapy's thin L3/L4 orchestration for a *fresh* System
(no cached neighbour list), parameterised by a kernel backend ``K`` which is
either ``oracle.ref`` (the reference C++ compiled unmodified) or
``oracle.port`` (our plain-C restatement).  It exists so the oracle can be
pinned against the reference's own golden fixtures
(tests/fixtures/structure_analysis/*.npz upstream; tests/golden/ here).

Follows: box.py:465-502 (thickness, check_small_box), tool_function.py:141-192
(replicate), neighbor.py:94-142, knn.py:79-129, system.py:1605-1636 (AJA),
1716-1861 (Steinhardt), 1926-1970 (PTM), 1986-2003 (CSP), 2030-2064 (CNA),
2273-2361 (RDF), common_neighbor_analysis.py:89-154,
polyhedral_template_matching.py:119-167,
radial_distribution_function.py:147-211.
"""
from __future__ import annotations

import numpy as np


def thickness(box):
    """box.py:465-481 get_thickness (NumPy arithmetic, as the reference's Python side)."""
    box = np.asarray(box, float).reshape(3, 3)
    vol = np.linalg.det(box)
    return np.array([
        vol / np.linalg.norm(np.cross(box[1], box[2])),
        vol / np.linalg.norm(np.cross(box[0], box[2])),
        vol / np.linalg.norm(np.cross(box[0], box[1])),
    ])


def check_small_box(box, boundary, rc):
    """box.py:483-502."""
    t = thickness(box)
    rep = np.ones(3, np.int32)
    for i in range(3):
        if boundary[i] == 1 and t[i] < 2 * rc:
            rep[i] = int(np.ceil(2.0 * rc / t[i]))
    return rep


def safe_repeat(box, boundary, safe_L=15):
    """system.py:2032-2036 / common_neighbor_analysis.py:104-107."""
    rep = np.ceil(safe_L / thickness(box)).astype(int)
    for i in range(3):
        if boundary[i] == 0:
            rep[i] = 1
    return rep


class Frame:
    """Positions + box of one configuration (SoA, like the polars columns)."""

    def __init__(self, pos, box, boundary=(1, 1, 1), origin=(0.0, 0.0, 0.0)):
        pos = np.asarray(pos, float)
        self.x = np.ascontiguousarray(pos[:, 0])
        self.y = np.ascontiguousarray(pos[:, 1])
        self.z = np.ascontiguousarray(pos[:, 2])
        self.box = np.ascontiguousarray(np.asarray(box, float).reshape(3, 3))
        self.boundary = np.asarray(boundary, np.int32)
        self.origin = np.asarray(origin, float)
        self.N = pos.shape[0]

    @property
    def pos(self):
        return np.stack([self.x, self.y, self.z], axis=1)

    def geom(self):
        return self.x, self.y, self.z, self.box, self.origin, self.boundary

    def replicate(self, K, nx, ny, nz):
        """tool_function.py:141-192."""
        new_pos = K.repeat_cell(self.box, self.pos, int(nx), int(ny), int(nz))
        new_box = self.box * np.array([nx, ny, nz]).reshape(3, 1)
        return Frame(new_pos, new_box, self.boundary, self.origin)


def neighbor(K, fr: Frame, rc, max_neigh=None):
    """neighbor.py:94-142 -> (frame_used, verlet, dist, nn)."""
    rep = check_small_box(fr.box, fr.boundary, rc)
    if rep.sum() != 3:
        fr = fr.replicate(K, *rep)
    if max_neigh is None:
        v, d, n = K.build_neighbor_auto(*fr.geom(), rc)
    else:
        v, d, n = K.build_neighbor(*fr.geom(), rc, max_neigh)
        if int(n.max(initial=0)) > max_neigh:
            raise ValueError("max_neigh is too small")
    return fr, v, d, n


def nearest(K, fr: Frame, k):
    """knn.py:63-129 -> (frame_used, idx, dist)."""
    rep = [1, 1, 1]
    if k > fr.N:
        assert fr.boundary.sum() > 0
        while np.prod(rep) * fr.N < k:
            for i in range(3):
                if fr.boundary[i] == 1:
                    rep[i] += 3
    if sum(rep) != 3:
        fr = fr.replicate(K, *rep)
    idx, dst = K.knn(*fr.geom(), k)
    return fr, idx, dst


def cal_cna(K, fr: Frame, rc=None):
    """system.py:2005-2064 + common_neighbor_analysis.py:81-154, fresh system."""
    N = fr.N
    if fr.boundary.sum() == 0 and N <= 14:
        return np.zeros(N, np.int32)
    rep = safe_repeat(fr.box, fr.boundary)
    if rep.sum() == 3 and rc is not None:
        f2, v, d, n = neighbor(K, fr, rc)
        return K.fcna(*f2.geom(), v, n, rc)[:N]
    f2 = fr
    if rep.sum() != 3:
        f2 = fr.replicate(K, *rep)
    if rc is None:
        f3, idx, _ = nearest(K, f2, 14)
        return K.acna(*f3.geom(), idx)[:N]
    rep2 = check_small_box(f2.box, f2.boundary, rc)
    if rep2.sum() != 3:
        f2 = f2.replicate(K, *rep2)
    f3, v, d, n = neighbor(K, f2, rc)
    return K.fcna(*f3.geom(), v, n, rc)[:N]


def cal_ids(K, fr: Frame):
    """system.py:1493-1529 + identify_diamond_structure.py:75-124, fresh system (kNN path, safe_L = 15)."""
    N = fr.N
    if fr.boundary.sum() == 0 and N <= 4:
        return np.zeros(N, np.int32)
    rep = safe_repeat(fr.box, fr.boundary, safe_L=15)
    f2 = fr.replicate(K, *rep) if rep.sum() != 3 else fr
    f3, idx, _ = nearest(K, f2, 4)
    return K.ids(*f3.geom(), idx)[:N]


def cal_csp(K, fr: Frame, nnei):
    """system.py:1972-2003, fresh system (kNN path)."""
    if fr.N <= nnei and fr.boundary.sum() == 0:
        return np.full(fr.N, 10000, float)
    f2, idx, _ = nearest(K, fr, nnei)
    return K.csp(*f2.geom(), idx, nnei)[: fr.N]


def cal_aja(K, fr: Frame):
    """system.py:1605-1636, fresh system (kNN path)."""
    if fr.N < 14 and fr.boundary.sum() == 0:
        return np.zeros(fr.N, np.int32)
    f2, idx, dst = nearest(K, fr, 14)
    return K.aja(*f2.geom(), idx, dst)[: fr.N]


def cal_ptm(K, fr: Frame, structure="fcc-hcp-bcc", rmsd_threshold=0.1, types=None):
    """system.py:1863-1970 + polyhedral_template_matching.py:110-167 -> (output, ptm_indices)."""
    N = fr.N
    if fr.boundary.sum() == 0 and N <= 18:
        return np.zeros((N, 8)), np.zeros((N, 18), np.int32)
    rep = safe_repeat(fr.box, fr.boundary)
    f2 = fr
    if rep.sum() != 3:
        f2 = fr.replicate(K, *rep)
        if types is not None:
            types = np.tile(types, int(np.prod(rep)))
    f3, idx, _ = nearest(K, f2, 18)
    t = np.ones(f3.N, np.int32) if types is None else np.asarray(types, np.int32)
    out, ind = K.ptm(structure, *f3.geom(), idx, t, rmsd_threshold)
    return out[:N], ind[:N]


def cal_steinhardt(K, fr: Frame, llist, nnn=0, rc=-1.0, average=False, wl=False, wlhat=False,
                   identify_liquid=False, threshold=0.7, n_bond=7):
    """system.py:1716-1861 + steinhardt_bond_orientation.py:203-302, fresh system."""
    if nnn > 0:
        f2, v, d = nearest(K, fr, nnn)
        n = np.full(f2.N, nnn, np.int32)
    else:
        assert rc > 0
        f2, v, d, n = neighbor(K, fr, rc)
    ll = np.asarray(llist, int)
    qn, qr, qi = K.get_sq(*f2.geom(), v, d, n, ll, nnn=nnn, rc=rc, average=average, wl=wl, wlhat=wlhat)
    res = {"qnarray": qn[: fr.N], "qlm_r": qr, "qlm_i": qi}
    if identify_liquid:
        q6i = int(np.where(ll == 6)[0][0])
        sl, nb = K.solid_liquid(q6i, np.ascontiguousarray(qn[:, q6i]), v, d, n, qr, qi, float(threshold),
                                int(n_bond), nnn=nnn, rc=rc)
        res["solidliquid"] = sl[: fr.N]
        res["nbond"] = nb[: fr.N]
    return res


def rdf_normalise(counts, type_list, ntype, rc, nbin, volume, N):
    """radial_distribution_function.py:147-211 -> (r, g_total, g_partial dict keyed by (a,b))."""
    edges = np.linspace(0, rc, nbin + 1)
    const = (4.0 * np.pi / 3.0 * (edges[1:] ** 3 - edges[:-1] ** 3)) / volume
    r = (edges[1:] + edges[:-1]) / 2
    number_per_type = np.bincount(type_list, minlength=ntype)
    total = np.zeros(nbin)
    for a in range(ntype):
        for b in range(ntype):
            total += counts[a, b]
    g_total = total / const / N**2
    part = {}
    for a in range(ntype):
        for b in range(a, ntype):
            raw = counts[a, b] if a == b else counts[a, b] + counts[b, a]
            if number_per_type[a] > 0 and number_per_type[b] > 0:
                g = raw / (number_per_type[a] * number_per_type[b]) / const
                if a != b:
                    g = g * 0.5
            else:
                g = np.zeros_like(r)
            part[(a, b)] = g
    return r, g_total, part


def cal_rdf(K, fr: Frame, rc, nbin, type_list=None, streaming=None):
    """system.py:2235-2361, fresh system -> dict(r, g_total, g_partial, counts)."""
    labels = np.zeros(fr.N, np.int32) if type_list is None else np.asarray(type_list)
    if streaming is None:
        t = thickness(fr.box)
        per = [t[i] for i in range(3) if fr.boundary[i]]
        streaming = rc >= (min(per) if per else float("inf")) / 3.0
    uniq = sorted(set(labels.tolist()))
    remap = {v: i for i, v in enumerate(uniq)}
    ntype = len(uniq)
    if streaming:
        rep = check_small_box(fr.box, fr.boundary, rc)
        f2 = fr
        if rep.sum() != 3:
            f2 = fr.replicate(K, *rep)
            labels = np.tile(labels, int(np.prod(rep)))
        tl = np.array([remap[v] for v in labels.tolist()], np.int32)
        counts = K.rdf_streaming(f2.x, f2.y, f2.z, tl, ntype, f2.box, f2.origin, f2.boundary, rc, nbin)
    else:
        f2, v, d, n = neighbor(K, fr, rc)
        if f2.N != fr.N:
            labels = np.tile(labels, f2.N // fr.N)
        tl = np.array([remap[v_] for v_ in labels.tolist()], np.int32)
        if ntype > 1:
            counts = K.rdf_list(v, d, n, tl, ntype, rc, nbin)
        else:
            counts = np.zeros((1, 1, nbin))
            counts[0, 0] = K.rdf_single(v, d, n, rc, nbin)
    vol = np.linalg.det(f2.box)
    r, g_total, part = rdf_normalise(counts, tl, ntype, rc, nbin, vol, f2.N)
    return {"r": r, "g_total": g_total, "g_partial": part, "counts": counts, "elements": uniq}


# --------------------------------------------------------------------------
# further list consumers
# --------------------------------------------------------------------------
def cal_cnp(K, fr: Frame, rc):
    """system.py:1572-1603: cut-off list (replicated for small boxes), compute_cnp, slice [:N]."""
    fu, v, d, n = neighbor(K, fr, rc)
    return K.cnp(*fu.geom(), v, d, n, rc)[: fr.N]


def cal_average_by_neighbor(K, fr: Frame, rc, value, include_self=True):
    """system.py:2363-2414 (big boxes only: no replica)."""
    fu, v, d, n = neighbor(K, fr, rc)
    assert fu.N == fr.N
    return K.average_by_neighbor(rc, v, d, n, value, include_self)


def cal_wcp(K, fr: Frame, rc, type_list, ntype):
    """system.py:1638-1676."""
    fu, v, d, n = neighbor(K, fr, rc)
    assert fu.N == fr.N
    return K.wcp(v, n, type_list, ntype)


def cal_structure_entropy(K, fr: Frame, rc, sigma, use_local_density=False, average_rc=0.0):
    """system.py:2480-2542: list on the (possibly replicated) frame, its volume, optional neighbour average."""
    fu, v, d, n = neighbor(K, fr, rc)
    vol = abs(float(np.linalg.det(fu.box)))
    ent = K.structure_entropy(rc, sigma, use_local_density, vol, d, n)
    if average_rc > 0:
        return K.average_by_neighbor(average_rc, v, d, n, ent, True)[: fr.N]
    return ent[: fr.N]
