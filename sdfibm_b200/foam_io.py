"""Readers for the binary OpenFOAM files the reference ships as fixtures.

Layout: ASCII ``FoamFile{...}`` header, then ``<count>\\n(`` + raw little-endian
payload + ``)``.  ``points`` = count x 3 fp64, ``owner``/``neighbour`` = count int32, ``faces`` is a
faceCompactList = offsets labelList (nFaces+1) followed by a flat labelList of vertex ids, and a
volScalarField is ``internalField nonuniform List<scalar> <count>(raw fp64)``.

Only the stand-alone harness and the fixture generator use this; in the drop-in OpenFOAM itself
owns the mesh (reference src/meshinfo.h:20-29).
"""
from __future__ import annotations

import os
import re

import numpy as np


def _find_list(buf: bytes, start: int = 0):
    """Return (count, payload_offset) of the first ``<count>\\n(`` after ``start``."""
    m = re.compile(rb"\n(\d+)\s*\n?\(").search(buf, start)
    if m is None:
        raise ValueError("no binary list found")
    return int(m.group(1)), m.end()


def _header_end(buf: bytes) -> int:
    i = buf.find(b"FoamFile")
    if i < 0:
        raise ValueError("not a FoamFile")
    return buf.find(b"}", i) + 1


def read_points(path: str) -> np.ndarray:
    buf = open(path, "rb").read()
    n, off = _find_list(buf, _header_end(buf))
    return np.frombuffer(buf, dtype="<f8", count=3 * n, offset=off).reshape(n, 3).copy()


def read_labels(path: str) -> np.ndarray:
    buf = open(path, "rb").read()
    n, off = _find_list(buf, _header_end(buf))
    return np.frombuffer(buf, dtype="<i4", count=n, offset=off).copy()


def read_faces(path: str):
    """faceCompactList -> (offsets[nFaces+1], flat vertex ids)."""
    buf = open(path, "rb").read()
    n, off = _find_list(buf, _header_end(buf))
    offsets = np.frombuffer(buf, dtype="<i4", count=n, offset=off).copy()
    m, off2 = _find_list(buf, off + 4 * n)
    ids = np.frombuffer(buf, dtype="<i4", count=m, offset=off2).copy()
    assert offsets[-1] == m
    return offsets, ids


def read_scalar_field(path: str) -> np.ndarray:
    buf = open(path, "rb").read()
    i = buf.find(b"internalField")
    n, off = _find_list(buf, i)
    return np.frombuffer(buf, dtype="<f8", count=n, offset=off).copy()


def read_polymesh(case_dir: str):
    """Return dict(points, face_off, face_pts, owner, neighbour) of ``case_dir/constant/polyMesh``."""
    d = os.path.join(case_dir, "constant", "polyMesh")
    face_off, face_pts = read_faces(os.path.join(d, "faces"))
    return dict(
        points=read_points(os.path.join(d, "points")),
        face_off=face_off,
        face_pts=face_pts,
        owner=read_labels(os.path.join(d, "owner")),
        neighbour=read_labels(os.path.join(d, "neighbour")),
    )


# ---- writers (binary, the layout the readers above take) and the decomposed-case layout -----------------------------------------
def _header(cls: str, obj: str, location: str) -> bytes:
    return (f"FoamFile\n{{\n    version     2.0;\n    format      binary;\n    class       {cls};\n"
            f"    location    \"{location}\";\n    object      {obj};\n}}\n\n").encode()


def _write_list(path: str, cls: str, obj: str, location: str, payloads) -> None:
    os.makedirs(os.path.dirname(path), exist_ok=True)
    with open(path, "wb") as f:
        f.write(_header(cls, obj, location))
        for count, raw in payloads:
            f.write(f"\n{count}\n(".encode() + raw + b")\n")


def write_polymesh(case_dir: str, points, face_off, face_pts, owner, neighbour, location: str = "constant/polyMesh") -> None:
    """``case_dir/constant/polyMesh/{points,faces,owner,neighbour}`` in OpenFOAM's binary format (faces as a faceCompactList)."""
    d = os.path.join(case_dir, "constant", "polyMesh")
    points = np.ascontiguousarray(points, dtype="<f8")
    face_off, face_pts = np.ascontiguousarray(face_off, dtype="<i4"), np.ascontiguousarray(face_pts, dtype="<i4")
    owner, neighbour = np.ascontiguousarray(owner, dtype="<i4"), np.ascontiguousarray(neighbour, dtype="<i4")
    _write_list(os.path.join(d, "points"), "vectorField", "points", location, [(len(points), points.tobytes())])
    _write_list(os.path.join(d, "faces"), "faceCompactList", "faces", location,
                [(len(face_off), face_off.tobytes()), (len(face_pts), face_pts.tobytes())])
    _write_list(os.path.join(d, "owner"), "labelList", "owner", location, [(len(owner), owner.tobytes())])
    _write_list(os.path.join(d, "neighbour"), "labelList", "neighbour", location, [(len(neighbour), neighbour.tobytes())])


def write_labels(path: str, obj: str, labels, location: str = "constant/polyMesh") -> None:
    labels = np.ascontiguousarray(labels, dtype="<i4")
    _write_list(path, "labelIOList", obj, location, [(len(labels), labels.tobytes())])


def write_scalar_field(path: str, obj: str, values) -> None:
    """A volScalarField with a nonuniform binary internalField and no boundary entries (what read_scalar_field takes)."""
    values = np.ascontiguousarray(values, dtype="<f8")
    os.makedirs(os.path.dirname(path), exist_ok=True)
    with open(path, "wb") as f:
        f.write(_header("volScalarField", obj, os.path.basename(os.path.dirname(path))))
        f.write(b"dimensions      [0 0 0 0 0 0 0];\n\ninternalField   nonuniform List<scalar> ")
        f.write(f"\n{len(values)}\n(".encode() + values.tobytes() + b");\n\nboundaryField\n{\n}\n")


def read_decomposed(case_dir: str):
    """A `decomposePar` case: one entry per ``processorK`` directory, in rank order, with the subdomain's polyMesh in
    LOCAL numbering and its ``cellProcAddressing`` (local cell -> global cell), which is all the multi-GPU path needs: cells are
    partitioned, solids replicated, no halo (every per-cell result depends only on that cell's geometry and U)."""
    ranks = []
    k = 0
    while os.path.isdir(os.path.join(case_dir, f"processor{k}")):
        d = os.path.join(case_dir, f"processor{k}")
        pm = read_polymesh(d)
        pm["cell_addressing"] = read_labels(os.path.join(d, "constant", "polyMesh", "cellProcAddressing"))
        ranks.append(pm)
        k += 1
    if not ranks:
        raise FileNotFoundError(f"no processor0 directory under {case_dir}")
    return ranks


def reconstruct_cells(ranks, fields, n_cells_total: int) -> np.ndarray:
    """reconstructPar for cell data: scatter each rank's field through its cellProcAddressing into the global numbering."""
    first = np.asarray(fields[0])
    out = np.zeros((n_cells_total,) + first.shape[1:], dtype=first.dtype)
    for pm, f in zip(ranks, fields):
        out[pm["cell_addressing"]] = f
    return out
