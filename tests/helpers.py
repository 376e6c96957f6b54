"""Synthetic configurations shared by the tests (NumPy only).

Lattices follow the reference generator bit for bit:
build_lattice.py:38-61 bases, repeat_cell.cpp:41-59 ordering
(cell-major, iz fastest; pos = basis*a + (ix*a, iy*a, iz*a)).
"""
import numpy as np

FCC = np.array([[0.0, 0.0, 0.0], [0.5, 0.5, 0.0], [0.0, 0.5, 0.5], [0.5, 0.0, 0.5]])
BCC = np.array([[0.0, 0.0, 0.0], [0.5, 0.5, 0.5]])


def lattice(basis, a, nx, ny, nz):
    old = (basis @ (a * np.eye(3)))
    ix, iy, iz = np.meshgrid(np.arange(nx), np.arange(ny), np.arange(nz), indexing="ij")
    shift = np.stack([ix.ravel() * a + iy.ravel() * 0.0 + iz.ravel() * 0.0,
                      ix.ravel() * 0.0 + iy.ravel() * a + iz.ravel() * 0.0,
                      ix.ravel() * 0.0 + iy.ravel() * 0.0 + iz.ravel() * a], axis=1)
    pos = (old[None, :, :] + shift[:, None, :]).reshape(-1, 3)
    box = np.diag([nx * a, ny * a, nz * a]).astype(float)
    return np.ascontiguousarray(pos), box


def fcc(a=3.615, n=8):
    return lattice(FCC, a, n, n, n)


def bcc(a=2.8665, n=8):
    return lattice(BCC, a, n, n, n)


def rattle(pos, sigma, seed):
    return pos + np.random.default_rng(seed).normal(0.0, sigma, pos.shape)


def shear(pos, box, xy=0.0, xz=0.0, yz=0.0):
    """Affine shear into a triclinic cell (rows are lattice vectors)."""
    F = np.array([[1.0, 0.0, 0.0], [xy, 1.0, 0.0], [xz, yz, 1.0]])
    return pos @ F, box @ F


def random_gas(n, L, seed):
    rng = np.random.default_rng(seed)
    return rng.random((n, 3)) * L, np.diag([L, L, L]).astype(float)


def same_rows_as_sets(v1, n1, v2, n2):
    """True when every row holds the same index multiset (ordering ignored)."""
    if not np.array_equal(n1, n2):
        return False
    for i in range(v1.shape[0]):
        k = min(int(n1[i]), v1.shape[1])
        if sorted(v1[i, :k].tolist()) != sorted(v2[i, :k].tolist()):
            return False
    return True


DIAMOND = np.concatenate([FCC, FCC + 0.25])


def diamond(a=3.567, n=5):
    return lattice(DIAMOND, a, n, n, n)


def hex_diamond(a=2.52, nx=6, ny=4, nz=4):
    """Lonsdaleite in its 8-atom orthogonal cell (a, sqrt(3) a, c = sqrt(8/3) a)."""
    c = np.sqrt(8.0 / 3.0) * a
    frac = []
    for (u, v) in ((0.0, 0.0), (0.5, 0.5)):            # the two hexagonal cells of the orthogonal cell
        for (fx, fy, fz) in ((0.0, 1 / 3, 0.0), (0.5, 1 / 6, 0.5), (0.0, 1 / 3, 3 / 8), (0.5, 1 / 6, 7 / 8)):
            frac.append(((fx + u) % 1.0, (fy + v) % 1.0, fz))
    frac = np.array(frac)
    cell = np.array([a, np.sqrt(3.0) * a, c])
    ix, iy, iz = np.meshgrid(np.arange(nx), np.arange(ny), np.arange(nz), indexing="ij")
    shift = np.stack([ix.ravel(), iy.ravel(), iz.ravel()], axis=1).astype(float)
    pos = ((frac[None, :, :] + shift[:, None, :]) * cell).reshape(-1, 3)
    return np.ascontiguousarray(pos), np.diag(cell * np.array([nx, ny, nz]))
