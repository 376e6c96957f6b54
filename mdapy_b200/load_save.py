"""Frame readers / writers feeding the device path, mirroring the text formats of the
reference's ``mdapy.load_save.BuildSystem`` (src/mdapy/load_save.py): LAMMPS dump (``read_dump`` 1337-1376,
frame parser 66-199, ``write_dump`` 1911-1990) and classical / extended XYZ (``read_xyz`` 653-862,
``write_xyz`` 1566-1653), ``.gz`` transparently (23-40).  Same column names, dtypes (int32 for
id/type/ix/iy/iz, float64 otherwise, str for element), box conventions (rows = cell vectors, row 4 = origin;
restricted-triclinic ``xy xz yz`` bounds converted like 109-126) and boundary flags (``pp`` = periodic).

Host-side text parsing only: the arrays land in pinned-friendly contiguous NumPy columns that ``System``
uploads once.  mdapy's binary ``.mp`` (parquet + key-value metadata, 610-650 / 1534-1564) goes through pyarrow.
LAMMPS data files and POSCAR are not on the hot path.
"""
from __future__ import annotations

import gzip
import io
import re
from typing import Any, Dict, List, Optional, Tuple

import numpy as np

from .box import Box
from .frame import Frame

INT_COLS = {"id", "type", "ix", "iy", "iz", "mol", "proc", "procp1"}
STR_COLS = {"element", "typelabel"}


def _open(filename: str, mode: str = "r"):
    if str(filename).endswith(".gz"):
        return gzip.open(filename, mode + "t")
    return open(filename, mode)


def _table(rows: List[str], names: List[str], str_cols=STR_COLS, int_cols=INT_COLS) -> Dict[str, np.ndarray]:
    """Whitespace-separated rows -> typed columns (pandas' C parser when available, NumPy otherwise)."""
    text = "".join(r if r.endswith("\n") else r + "\n" for r in rows)
    ncol = len(names)
    try:
        import pandas as pd

        df = pd.read_csv(io.StringIO(text), sep=r"\s+", header=None, names=list(range(ncol)), usecols=list(range(ncol)),
                         dtype={j: (str if names[j] in str_cols else np.float64) for j in range(ncol)}, engine="c", float_precision="round_trip")
        cols = {names[j]: df[j].to_numpy() for j in range(ncol)}
    except ImportError:  # pragma: no cover
        cells = np.array([r.split()[:ncol] for r in rows])
        cols = {names[j]: cells[:, j] for j in range(ncol)}
    out: Dict[str, np.ndarray] = {}
    for name in names:
        c = cols[name]
        if name in str_cols:
            out[name] = np.asarray(c, dtype=object)
        elif name in int_cols:
            out[name] = np.ascontiguousarray(np.asarray(c, dtype=np.float64).astype(np.int32))
        else:
            out[name] = np.ascontiguousarray(np.asarray(c, dtype=np.float64))
    return out


# --------------------------------------------------------------------------- LAMMPS dump
def parse_dump_frame(lines: List[str], source: str = "<dump>") -> Tuple[Frame, Box, Dict[str, Any]]:
    """One dump frame: 9 header lines + N atom rows (load_save.py:66-199)."""
    if len(lines) < 9:
        raise ValueError(f"{source}: dump frame has only {len(lines)} lines (<9)")
    try:
        timestep = int(lines[1].strip())
    except (IndexError, ValueError):
        raise ValueError(f"{source}: malformed ITEM: TIMESTEP value")
    try:
        n_atoms = int(lines[3].strip())
    except (IndexError, ValueError):
        raise ValueError(f"{source}: malformed ITEM: NUMBER OF ATOMS value")
    bb_line = lines[4].strip()
    if not bb_line.startswith("ITEM: BOX BOUNDS"):
        raise ValueError(f"{source}: expected 'ITEM: BOX BOUNDS' on line 5")
    tokens = bb_line.split()[3:]
    if tokens and all(t in {"pp", "ff", "ss", "mm"} for t in tokens[-3:]):
        boundary = [1 if t == "pp" else 0 for t in tokens[-3:]]
        geometry = tokens[:-3]
    else:
        boundary, geometry = [1, 1, 1], tokens
    rows = [lines[5].split(), lines[6].split(), lines[7].split()]
    if "abc" in geometry and "origin" in geometry:
        box = np.vstack([np.array(rows[k][:3], dtype=np.float64) for k in range(3)] +
                        [np.array([rows[0][3], rows[1][3], rows[2][3]], dtype=np.float64)])
    elif {"xy", "xz", "yz"}.issubset(geometry):
        xlo_b, xhi_b, xy = (float(v) for v in rows[0][:3])
        ylo_b, yhi_b, xz = (float(v) for v in rows[1][:3])
        zlo_b, zhi_b, yz = (float(v) for v in rows[2][:3])
        xlo = xlo_b - min(0.0, xy, xz, xy + xz)
        xhi = xhi_b - max(0.0, xy, xz, xy + xz)
        ylo = ylo_b - min(0.0, yz)
        yhi = yhi_b - max(0.0, yz)
        box = np.array([[xhi - xlo, 0, 0], [xy, yhi - ylo, 0], [xz, yz, zhi_b - zlo_b], [xlo, ylo, zlo_b]], float)
    else:
        lo = [float(rows[k][0]) for k in range(3)]
        hi = [float(rows[k][1]) for k in range(3)]
        box = np.array([[hi[0] - lo[0], 0, 0], [0, hi[1] - lo[1], 0], [0, 0, hi[2] - lo[2]], lo], float)
    header = lines[8].rstrip()
    if not header.startswith("ITEM: ATOMS"):
        raise ValueError(f"{source}: expected 'ITEM: ATOMS' on line 9")
    names = header.split()[2:]
    body = lines[9: 9 + n_atoms]
    if len(body) != n_atoms:
        raise ValueError(f"{source}: expected {n_atoms} atom rows, got {len(body)}")
    cols = _table(body, names) if n_atoms else {n: np.zeros(0) for n in names}
    have = set(cols)
    if not {"x", "y", "z"}.issubset(have):
        for tag in ("xs", "xsu"):
            ty, tz = tag.replace("x", "y"), tag.replace("x", "z")
            if {tag, ty, tz}.issubset(have):
                scaled = np.stack([cols.pop(tag), cols.pop(ty), cols.pop(tz)], axis=1)
                absolute = box[3] + scaled @ box[:3]
                cols.update(x=absolute[:, 0].copy(), y=absolute[:, 1].copy(), z=absolute[:, 2].copy())
                break
        else:
            if {"xu", "yu", "zu"}.issubset(have):
                cols["x"], cols["y"], cols["z"] = cols.pop("xu"), cols.pop("yu"), cols.pop("zu")
    return Frame(cols), Box(box[:3], boundary, box[3]), {"timestep": timestep}


def read_dump(filename: str) -> Tuple[Frame, Box, Dict[str, Any]]:
    """First frame of a LAMMPS text dump (``.dump`` / ``.dump.gz``)."""
    with _open(filename) as f:
        head = [f.readline() for _ in range(4)]
        try:
            n = int(head[3].strip())
        except ValueError:
            raise ValueError(f"{filename}: malformed ITEM: NUMBER OF ATOMS value")
        rest = [f.readline() for _ in range(5 + n)]
    return parse_dump_frame(head + rest, str(filename))


def iter_dump_frames(filename: str):
    """Every frame of a multi-frame dump, one (Frame, Box, info) at a time -- the frame-streaming entry
    point for trajectory analysis (one ``DeviceSystem`` handle can be reused across the frames)."""
    with _open(filename) as f:
        while True:
            head = [f.readline() for _ in range(4)]
            if not head[0]:
                return
            n = int(head[3].strip())
            rest = [f.readline() for _ in range(5 + n)]
            yield parse_dump_frame(head + rest, str(filename))


def write_dump(filename: str, box: Box, data: Frame, timestep: int = 0, columns: Optional[List[str]] = None) -> None:
    cols = columns or [c for c in data.columns if c not in STR_COLS or c == "element"]
    b = [[float(v) for v in row] for row in np.asarray(box.box, float)]     # plain floats: repr() round-trips
    o = [float(v) for v in np.asarray(box.origin, float)]
    flags = " ".join("pp" if v else "ff" for v in box.boundary)
    with _open(filename, "w") as f:
        f.write(f"ITEM: TIMESTEP\n{int(timestep)}\nITEM: NUMBER OF ATOMS\n{data.shape[0]}\n")
        ortho = abs(b[0][1]) + abs(b[0][2]) + abs(b[1][0]) + abs(b[1][2]) + abs(b[2][0]) + abs(b[2][1]) < 1e-12
        if ortho:
            f.write(f"ITEM: BOX BOUNDS {flags}\n")
            for k in range(3):
                f.write(f"{o[k]!r} {o[k] + b[k][k]!r}\n")
        elif abs(b[0][1]) + abs(b[0][2]) + abs(b[1][2]) < 1e-12:      # restricted triclinic
            xy, xz, yz = b[1][0], b[2][0], b[2][1]
            xlo_b = o[0] + min(0.0, xy, xz, xy + xz)
            xhi_b = o[0] + b[0][0] + max(0.0, xy, xz, xy + xz)
            ylo_b = o[1] + min(0.0, yz)
            yhi_b = o[1] + b[1][1] + max(0.0, yz)
            f.write(f"ITEM: BOX BOUNDS xy xz yz {flags}\n")
            f.write(f"{xlo_b!r} {xhi_b!r} {xy!r}\n{ylo_b!r} {yhi_b!r} {xz!r}\n{o[2]!r} {o[2] + b[2][2]!r} {yz!r}\n")
        else:                                                           # general triclinic
            f.write(f"ITEM: BOX BOUNDS abc origin {flags}\n")
            for k in range(3):
                f.write(f"{b[k][0]!r} {b[k][1]!r} {b[k][2]!r} {o[k]!r}\n")
        f.write("ITEM: ATOMS " + " ".join(cols) + "\n")
        arrays = [np.asarray(data[c]) for c in cols]
        for row in zip(*arrays):
            f.write(" ".join(str(int(v)) if isinstance(v, (np.integer, int)) else (v if isinstance(v, str) else repr(float(v)))
                             for v in row) + "\n")


# --------------------------------------------------------------------------- XYZ
_XYZ_ALIASES = {"pos": ["x", "y", "z"], "unwrapped_position": ["xu", "yu", "zu"], "unwrapped_pos": ["xu", "yu", "zu"],
                "vel": ["vx", "vy", "vz"], "velo": ["vx", "vy", "vz"], "forces": ["fx", "fy", "fz"],
                "force": ["fx", "fy", "fz"]}


def read_xyz(filename: str) -> Tuple[Frame, Box, Dict[str, Any]]:
    """Classical or extended XYZ (load_save.py:653-862).  A classical file has no cell: the box is the
    bounding box of the positions with open boundaries, as the reference does."""
    with _open(filename) as f:
        lines = f.readlines()
    if len(lines) < 2:
        raise ValueError(f"{filename}: too short to be an XYZ file")
    natom = int(lines[0].strip())
    if natom < 0:
        raise ValueError(f"{filename}: negative atom count {natom}")
    if len(lines) < 2 + natom:
        raise ValueError(f"{filename}: header says {natom} atoms but only {len(lines) - 2} body lines present")
    comment = lines[1].rstrip("\r\n")
    info: Dict[str, Any] = {}
    for m in re.findall(r'(\w+)=(?:"([^"]+)"|([^ ]+))', comment.replace("'", '"')):
        info[m[0].lower()] = m[1] if m[1] else m[2]
    names: List[str] = []
    kinds: Dict[str, str] = {}
    if "properties" in info:
        content = info["properties"].strip().split(":")
        i = 0
        while i + 2 < len(content):
            name, ptype, ncol = content[i], content[i + 1], int(content[i + 2])
            if ptype not in ("S", "R", "I"):
                raise ValueError(f"{filename}: unrecognised XYZ type {ptype!r}")
            if name in _XYZ_ALIASES and ptype == "R" and ncol == 3 and _XYZ_ALIASES[name][0] not in kinds:
                sub = _XYZ_ALIASES[name]
            elif name in ("species", "element") and ptype == "S" and ncol == 1 and "element" not in kinds:
                sub = ["element"]
            elif ncol == 1:
                sub = [name]
            else:
                sub = [f"{name}_{k}" for k in range(ncol)]
            for s_ in sub:
                names.append(s_)
                kinds[s_] = ptype
            i += 3
    else:
        names, kinds = ["element", "x", "y", "z"], {"element": "S", "x": "R", "y": "R", "z": "R"}
    body = lines[2: 2 + natom]
    text_cols = {n for n, k in kinds.items() if k == "S"}
    int_cols = {n for n, k in kinds.items() if k == "I"}
    cols = (_table(body, names, STR_COLS | text_cols, INT_COLS | int_cols) if natom
            else {n: np.zeros(0) for n in names})
    # pbc is honoured with or without a Lattice (load_save.py of the reference parses it first);
    # only "T" / "1" count as periodic, like the reference
    boundary = None
    if "pbc" in info:
        boundary = [1 if t in ("T", "1") else 0 for t in info["pbc"].split()]
    if "lattice" in info:
        cell = np.array(info["lattice"].split(), float).reshape(3, 3)
        origin = np.array(info["origin"].split(), float) if "origin" in info else np.zeros(3)
        box = Box(cell, boundary if boundary is not None else [1, 1, 1], origin)
    else:
        pos = np.stack([cols["x"], cols["y"], cols["z"]], axis=1)
        lo, hi = pos.min(axis=0), pos.max(axis=0)
        ext = hi - lo
        box = Box(np.diag(np.where(ext > 0, ext, 1e-9)), boundary if boundary is not None else [0, 0, 0], lo)
    for k in ("lattice", "properties", "pbc", "origin"):
        info.pop(k, None)
    return Frame(cols), box, info


def write_xyz(filename: str, box: Box, data: Frame) -> None:
    """Extended XYZ with Lattice / Origin / pbc and every numeric column."""
    cols = [c for c in data.columns if c not in ("x", "y", "z", "element")]
    props = ["species:S:1"] if "element" in data.columns else []
    props.append("pos:R:3")
    for c in cols:
        props.append(f"{c}:{'I' if np.issubdtype(np.asarray(data[c]).dtype, np.integer) else 'R'}:1")
    b, o = np.asarray(box.box, float), np.asarray(box.origin, float)
    lat = " ".join(repr(float(v)) for v in b.reshape(-1))
    org = " ".join(repr(float(v)) for v in o)
    pbc = " ".join("T" if v else "F" for v in box.boundary)
    order = (["element"] if "element" in data.columns else []) + ["x", "y", "z"] + cols
    arrays = [np.asarray(data[c]) for c in order]
    with _open(filename, "w") as f:
        f.write(f"{data.shape[0]}\n")
        f.write(f'Lattice="{lat}" Properties={":".join(props)} pbc="{pbc}" Origin="{org}"\n')
        for row in zip(*arrays):
            f.write(" ".join(v if isinstance(v, str) else (str(int(v)) if isinstance(v, (np.integer, int)) else repr(float(v)))
                             for v in row) + "\n")


# --------------------------------------------------------------------------- .mp (parquet)
def read_mp(filename: str) -> Tuple[Frame, Box, Dict[str, Any]]:
    """mdapy's native ``.mp`` file: a parquet table whose key-value metadata carries ``box`` (9 numbers),
    ``origin`` and ``boundary`` as strings (load_save.py:610-650).  Read with pyarrow (polars is not needed)."""
    import pyarrow.parquet as pq

    table = pq.read_table(filename)
    meta = {k.decode(): v.decode() for k, v in (table.schema.metadata or {}).items()}
    cols: Dict[str, np.ndarray] = {}
    for name in table.column_names:
        col = table.column(name).combine_chunks()
        a = col.to_numpy(zero_copy_only=False)
        if a.dtype.kind in "OUS":
            a = np.asarray(a, dtype=object)
        cols[name] = np.ascontiguousarray(a) if a.dtype != object else a
    if "box" in meta:
        cell = np.array(meta["box"].split(), float).reshape(3, 3)
    else:   # no stored cell: bounding box of the atom cloud, zero extents padded (610-637)
        ext = np.array([cols[c].max() - cols[c].min() for c in ("x", "y", "z")], float)
        cell = np.diag(np.where(ext > 0, ext, 1e-9))
    origin = np.array(meta["origin"].split(), float) if "origin" in meta else None
    boundary = np.array(meta["boundary"].split(), np.int32) if "boundary" in meta else None
    info = {k: v for k, v in meta.items() if k not in ("box", "origin", "boundary") and not k.startswith("ARROW")
            and k != "pandas"}
    return Frame(cols), Box(cell, boundary, origin), info


def write_mp(filename: str, box: Box, data: Frame, global_info: Optional[Dict[str, Any]] = None) -> None:
    """Write ``.mp`` (load_save.py:1534-1564): parquet + box / origin / boundary (+ energy, stress, virial,
    timestep) in the key-value metadata."""
    import pyarrow as pa
    import pyarrow.parquet as pq

    arrays, names = [], []
    for c in data.columns:
        a = np.asarray(data[c])
        arrays.append(pa.array(a.tolist() if a.dtype == object else a))
        names.append(c)
    meta = {"box": " ".join(np.asarray(box.box, float).astype(str).flatten().tolist()),
            "origin": " ".join(np.asarray(box.origin, float).astype(str).tolist()),
            "boundary": " ".join(np.asarray(box.boundary).astype(str).tolist())}
    for k, v in (global_info or {}).items():
        if k in ("energy", "stress", "virial", "timestep"):
            meta[str(k)] = str(v)
    table = pa.Table.from_arrays(arrays, names=names).replace_schema_metadata(meta)
    pq.write_table(table, filename)


def from_file(filename: str) -> Tuple[Frame, Box, Dict[str, Any]]:
    """Dispatch on the extension like BuildSystem.from_file (load_save.py:358-411)."""
    name = str(filename)
    base = name[:-3] if name.endswith(".gz") else name
    ext = base.rsplit(".", 1)[-1].lower() if "." in base else ""
    if ext in ("dump", "lammpstrj"):
        return read_dump(name)
    if ext == "xyz":
        return read_xyz(name)
    if ext == "mp":
        return read_mp(name)
    raise NotImplementedError(
        f"{filename}: LAMMPS dump (.dump), XYZ (.xyz) -- optionally .gz -- and mdapy's .mp (parquet) are read here; "
        "other formats are outside the hot path (SURVEY.md 2.2)")
