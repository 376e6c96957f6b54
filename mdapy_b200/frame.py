"""Column frame standing in for the polars DataFrame the reference keeps in ``System.data``.

polars is not installable in this image, so per-atom columns
live in an ordered dict of equal-length NumPy arrays with the handful of
DataFrame methods the hot-path host code uses (``columns``, ``shape``,
``__getitem__`` -> column with ``to_numpy``, ``with_columns``, ``select``).
If polars is importable, ``Frame.to_polars()`` / ``Frame.from_any`` convert.
"""
from __future__ import annotations

from typing import Dict

import numpy as np


class Column(np.ndarray):
    """ndarray that also answers ``to_numpy`` like a polars Series."""

    def to_numpy(self, allow_copy: bool = True, writable: bool = False):
        return np.asarray(self)


class Frame:
    def __init__(self, columns: Dict[str, np.ndarray] | None = None):
        self._cols: Dict[str, np.ndarray] = {}
        n = None
        for k, v in (columns or {}).items():
            a = np.ascontiguousarray(v)
            if n is None:
                n = a.shape[0]
            elif a.shape[0] != n:
                raise ValueError(f"column {k!r} has length {a.shape[0]}, expected {n}")
            self._cols[k] = a

    @classmethod
    def from_any(cls, data) -> "Frame":
        if isinstance(data, Frame):
            return Frame(dict(data._cols))
        if isinstance(data, dict):
            return Frame(data)
        if hasattr(data, "columns") and hasattr(data, "to_numpy"):  # polars / pandas
            return Frame({c: np.asarray(data[c]) for c in data.columns})
        raise TypeError(f"unsupported data container {type(data)}")

    @property
    def columns(self):
        return list(self._cols.keys())

    @property
    def shape(self):
        n = next(iter(self._cols.values())).shape[0] if self._cols else 0
        return (n, len(self._cols))

    def __contains__(self, k):
        return k in self._cols

    def __getitem__(self, k: str) -> Column:
        return self._cols[k].view(Column)

    def with_columns(self, **cols) -> "Frame":
        new = dict(self._cols)
        n = self.shape[0]
        for k, v in cols.items():
            a = np.asarray(v)
            if a.ndim == 0:
                a = np.full(n, a)
            new[k] = a
        return Frame(new)

    def select(self, *names) -> "Frame":
        return Frame({k: self._cols[k] for k in names})

    def to_numpy(self) -> np.ndarray:
        return np.stack([self._cols[k] for k in self._cols], axis=1)

    def to_dict(self):
        return dict(self._cols)

    def to_polars(self):
        import polars as pl  # optional

        return pl.DataFrame(self._cols)

    def __repr__(self):
        return f"Frame(shape={self.shape}, columns={self.columns})"
