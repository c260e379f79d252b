#!/usr/bin/env python
"""Timed lines for the kernels bench.py's contract line does not cover: k_fix_internal against its
8 nC + 48 #(Ct >= 4) streaming roofline, and the collision step (k_col_*) at 10^4 / 10^5 spheres with delta = 2 r_max.
One JSON line per measurement on stdout.  Device-resident entries, CUDA events on the library's stream."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def hbm_peak():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        return 6650.0


def main():
    import torch
    import __graft_entry__ as ge
    ge.build()
    from sdfibm_b200 import capi, cases
    from sdfibm_b200.context import Context

    n = int(os.environ.get("AUX_N", "256"))
    scale = n / 256.0
    n_side = max(1, int(round(22 * scale)))
    case = cases.case_c4(n=n, n_solids=min(n_side ** 3, int(round(10000 * scale ** 3))), n_side=n_side)
    mesh = case["mesh"]
    nC, nS = mesh.n_cells, len(case["solids"])
    dev = torch.device("cuda", 0)
    ctx = Context(0)
    ctx.set_mesh(mesh, False)
    ctx.set_shapes(case["shapes"])
    ext = torch.cuda.ExternalStream(ctx.stream_ptr(), device=dev)
    dU = torch.from_numpy(case["U"]).to(dev)
    f = [torch.empty(k, dtype=torch.float64, device=dev) for k in (nC, 3 * nC, nC, nC, 6 * nS)]
    solids = capi.pinned_like(np.ascontiguousarray(case["solids"], dtype=capi.SOLID_DTYPE))
    ctx.interact_device(solids, dU.data_ptr(), case["dt"], case["rhof"], *[x.data_ptr() for x in f])
    n_inside = int((f[3] >= 4).sum().item())

    def timed(fn, steps=20, warmup=3):
        with torch.cuda.stream(ext):
            for _ in range(warmup):
                fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(ext)
            for _ in range(steps):
                fn()
            e1.record(ext)
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) / steps

    peak = hbm_peak()
    # ---- fixInternal (solidcloud.cpp:288-301): U = v + omega x (cc - x) where Ct >= 4 ----
    kern = []

    def fix():
        ctx.fix_internal_device(solids, dU.data_ptr(), f[3].data_ptr())
        kern.append(ctx.last_aux_timings()["fix_internal_ms"])
    ms = timed(fix)
    k_ms = float(np.mean(kern[-20:]))     # the kernel alone (CUDA events around it inside the entry)
    alg = 8 * nC + 48 * n_inside          # Ct read everywhere; cell centre read + U written where Ct >= 4
    print(json.dumps({"kernel": "k_fix_internal", "workload": f"C4 {n}^3, {nS} spheres", "cells": nC, "cells_inside": n_inside,
                      "ms_kernel": k_ms, "ms_call": ms,
                      "note": "ms_kernel: k_fix_internal alone (CUDA events on the library's stream); ms_call: sdfibm_fix_internal_device = solid "
                              "records H2D (pinned) + the kernel + one host synchronisation",
                      "roofline": {"bound": "hbm", "algorithmic_bytes": alg, "achieved": alg / (k_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                                   "frac": alg / (k_ms * 1e-3) / 1e9 / peak, "frac_of_the_call": alg / (ms * 1e-3) / 1e9 / peak,
                                   "formula": "8 nC + 48 #(Ct >= 4)"}}), flush=True)
    # ---- apply forcing (main.cpp:70-77) ----
    dT = torch.full((nC,), 300.0, dtype=torch.float64, device=dev)
    touched = int(sum(ctx.candidate_counts()))

    def forcing():
        ctx.apply_forcing_device(dU.data_ptr(), dT.data_ptr(), case["dt"])
    ms = timed(forcing)
    tc = ctx.touched_cells()
    nT = len(tc["cells"])
    alg = 5 * nC + nT * (24 + 8 + 8 + 2 * 24 + 2 * 8)
    print(json.dumps({"kernel": "k_apply_forcing", "workload": f"C4 {n}^3, {nS} spheres", "cells": nC, "cells_touched": nT, "ms": ms,
                      "roofline": {"bound": "hbm", "algorithmic_bytes": alg, "achieved": alg / (ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                                   "frac": alg / (ms * 1e-3) / 1e9 / peak, "formula": "5 nC (n_item, orig) + 104 #touched (Fs, As, Ts read; U, T read + written)"}}), flush=True)
    ctx.close()
    # ---- collision step (solidcloud.cpp:477-519): UGrid broad phase + narrow phase, delta = 2 r_max ----
    for n_sph in (10_000, 100_000):
        L = 256.0 if n_sph == 10_000 else 512.0
        side = int(round(n_sph ** (1 / 3))) + 1
        rng = np.random.RandomState(7)
        r = 5.0 if n_sph == 10_000 else 4.5
        from sdfibm_b200.shapes import make_shape, make_solids
        from sdfibm_b200.mesh import Mesh
        m2 = Mesh.hex_block((8, 8, 8), (0, 0, 0), (L / 8, L / 8, L / 8))
        shapes = np.array([make_shape("Sphere", radius=r)])
        S = make_solids(n_sph)
        g = np.stack(np.meshgrid(*[np.arange(side)] * 3, indexing="ij"), -1).reshape(-1, 3)[:n_sph]
        S["pos"] = (g + 0.5) * (L / side) + rng.uniform(-0.12, 0.12, size=(n_sph, 3)) * (L / side)
        c2 = Context(0)
        c2.set_mesh(m2, False)
        c2.set_shapes(shapes)
        ft = np.zeros((n_sph, 6))
        t, td = [], []
        for _ in range(8):
            t0 = time.perf_counter()
            pairs, _ = c2.collide(S, 2.0 * r, ft, capacity=40 * n_sph)
            t.append((time.perf_counter() - t0) * 1e3)
            td.append(c2.last_aux_timings()["collide_ms"])
        print(json.dumps({"kernel": "k_col_keys + radix sort + k_col_pairs (count, scan, emit) + k_col_narrow", "solids": n_sph, "delta": 2.0 * r,
                          "pairs": int(len(pairs)), "ms_device": float(np.median(td[2:])), "ms_wall_host_buffers": float(np.median(t[2:])),
                          "note": "ms_device: CUDA events around the kernels (key build, radix sort, pair count + scan + emit, narrow phase, with the "
                                  "4-byte pair-count read-back in the middle); ms_wall: sdfibm_collide through the host-buffer C ABI (solid records "
                                  "H2D, pair list + force/torque D2H, python wrapper); work buffers persist in the context"}), flush=True)
        c2.close()


if __name__ == "__main__":
    main()
