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
