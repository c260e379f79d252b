#!/usr/bin/env python
"""bench.py — IBM cell-solid updates/s of SolidCloud::interact on B200 (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload auto|c4|c5] [--impl ours|reference]

A "step" is one interact() over the whole workload: N=1 -> C4 (256^3 hex cells, 10^4 spheres r=5, the
configuration the roofline target is quoted on); N>1 -> C5 (512^3, 10^5 spheres/ellipsoids) block-split
like `decomposePar simple`, one subdomain per GPU, solids replicated, one NCCL all-reduce of the
per-solid force/torque per step.  `value` = candidate-list entries (cell-solid updates) of all ranks per
second with every input resident in HBM.  `e2e` = the step a host with device-resident flow fields makes:
solid states H2D, sdfibm_interact_device + sdfibm_apply_forcing_device, force/torque D2H, one host synchronisation;
`e2e_host_fields` = the host-buffer C-ABI call (sdfibm_interact) with pinned host arrays, H2D of U and D2H of As/Fs/Ts/Ct
inside the timed region (what the unchanged OpenFOAM loop, whose fields live in host memory, pays).
One JSON line on stdout (rank 0).
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "IBM cell-solid updates/sec (interact())"
UNIT = "cell-solid updates/s"


def measured_hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smmax, reasons, power = [], [], set(), []
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smmax.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smmax) if smmax else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


def measured_traffic(workload):
    """DRAM bytes per launch of the interact kernels from the committed ncu capture of the current kernels
    (profiles/r02_traffic.json); None when the capture is of another workload."""
    p = os.path.join(ROOT, "profiles", "r02_traffic.json")
    try:
        d = json.load(open(p))
        if d.get("workload", "").lower() == workload:
            return int(sum(d["kernels"].values())), d
    except Exception:
        pass
    return None, None


def algorithmic_bytes(n_cells, counts, n_solids):
    """B = 48 nCells + 112 P + 216 P_b + 176 N  (hex mesh)."""
    P = sum(counts)
    Pb = counts[1] + counts[2]
    return 48 * n_cells + 112 * P + 216 * Pb + 176 * n_solids


L2_NOTE = "inputs larger than L2 (fields + mesh arrays of one step >> 126 MB); no explicit flush"


L2_NOTE_SMALL = ("the reference's small example set-up: one step's data fits the 126 MB L2 and is not flushed; the figure of interest is "
                 "latency (ms/step of one graph launch), no roofline claim is made on it")


def make_config(desc):
    """`config` of the JSON line: the same keys and values in both arms (ours and --impl reference)."""
    small = desc.get("workload", "").startswith(("C1", "C2", "C3"))
    return dict(desc, l2=L2_NOTE_SMALL if small else L2_NOTE)


def build_case(args, rank, world):
    from sdfibm_b200 import cases

    wl = args.workload
    if wl == "auto":
        wl = "c4" if world == 1 else "c5"
    if wl in ("c1", "c2", "c3", "c3b", "c3tc"):
        # BASELINE configs 1-3 (latency-bound; ms/step is the figure of interest): the reference's own example set-ups
        if world != 1:
            raise SystemExit(f"workload {wl} is single-GPU")
        if wl == "c1":
            case = cases.case_c1()
            desc = {"workload": "C1: flow_past_cylinder central block 120x120x1 (dx 0.1), 1 fixed Circle r=1 (2-D)"}
        elif wl == "c2":
            case = cases.case_c2(with_walls=True)
            desc = {"workload": "C2: sedimentation 400x400x1 (dx 0.02), 100 Circle r=0.15 + 4 wall planes (2-D)"}
        elif wl == "c3":
            pts = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "tests", "golden", "m2_points.npz"))["points"]
            case = cases.case_c3(pts)
            desc = {"workload": "C3: falling_ellipse on the shipped 100x200x1 mesh, Ellipse 0.3/0.15 at -45 deg (2-D)"}
        elif wl == "c3tc":
            case = cases.case_taylor_couette()
            desc = {"workload": "C3 taylor_couette: the five-block O-grid of examples/taylor_couette/system/blockMeshDict (5 x 50x50x1 curved hexahedra), "
                                "Circle r=0.3 spinning at the origin + a Circle_Tail across a block junction (2-D)"}
        else:
            case = cases.case_skewed_2d()
            desc = {"workload": "C3b: rotated + jittered 60x60x1 quad block, Circle r=0.3 + Circle_TwoTail"}
    elif wl == "c4":
        n = args.n or 256
        scale = n / 256.0
        n_side = max(1, int(round(22 * scale)))
        n_solids = args.solids or min(n_side ** 3, int(round(10000 * scale ** 3)))
        if world == 1:
            case = cases.case_c4(n=n, n_solids=n_solids, n_side=n_side)
        else:  # weak-scaled C4: one n^3 block per rank is not a named config; use the C5 split instead
            raise SystemExit("workload c4 is single-GPU; use --workload c5 for N>1")
        desc = {"workload": f"C4: {n}^3 hex cells, {n_solids} Sphere r=5 on a jittered {n_side}^3 lattice, seed 12345"}
    else:
        n = args.n or 512
        scale = n / 512.0
        n_side = max(1, int(round(47 * scale)))
        n_solids = args.solids or min(n_side ** 3, int(round(100000 * scale ** 3)))
        case = cases.case_c5_block(rank, world, n=n, n_solids=n_solids, n_side=n_side)
        desc = {"workload": f"C5: {n}^3 hex cells, {n_solids} Sphere r=4.5 / Ellipsoid (5,4.5,4) mixed, jittered {n_side}^3 lattice, "
                            f"seed 12345; block split {cases.decompose_simple(world)}",
                "parallelism": f"decomposePar-simple x{world}, solids replicated, 1 NCCL allreduce(6N fp64)/step inside the library (sdfibm_comm_init)"}
    return wl, case, desc


def representative_solids(case, k):
    """k solid indices that represent the rank's mix of solids that touch its block and solids that do not (the reference loops
    over ALL solids on every rank, and the two kinds cost very different amounts): the solids are ordered touching-first and
    sampled at even spacing, so the sample keeps their proportion (k = 1 takes a touching one)."""
    S, m = case["solids"], case["mesh"]
    rb = np.array([float(sh["radiusB"]) for sh in case["shapes"]])
    r = np.where(rb[S["shape"]] > 0, rb[S["shape"]], 0.0) + 1.0
    pos = S["pos"]
    touching = np.all((pos + r[:, None] >= m.bounds_min) & (pos - r[:, None] <= m.bounds_max), axis=1)
    order = np.concatenate([np.nonzero(touching)[0], np.nonzero(~touching)[0]])
    k = max(1, min(int(k), len(order)))
    return order[np.unique(np.linspace(0, len(order) - 1, k).astype(int))] if k > 1 else order[:1]


_REF_CACHE = {}


def reference_available():
    """oracle/_ref/libsdfibm_ref.so: the reference's own hot-path translation units, compiled unmodified where /root/reference
    exists (the .so travels to the GPU box with the snapshot)."""
    try:
        from oracle import ref_py
        return ref_py.available()
    except Exception:
        return False


def cpu_baseline_sample(case, n_sample, faithful=True, repeats=1, indices=None, prefer_reference=True):
    """Time the reference's CPU algorithm on a sample of the case's solids (single thread, like one reference rank): the first
    n_sample, or the given indices.  Returns (pairs, ms of the reference's own timed region, kind): kind "reference" = the
    reference's compiled CellEnumerator / GeometricTools / Solid / shape classes inside the loop of src/solidcloud.cpp:361-464
    (oracle/_ref), "port" = the oracle's restatement (also used for the non-faithful variant)."""
    from oracle.oracle_py import Oracle

    o = Oracle(case["mesh"], case["two_d"])
    solids = case["solids"] if indices is None else np.ascontiguousarray(case["solids"][indices])
    n_sample = min(n_sample, len(solids)) if indices is None else len(solids)
    best = None
    if faithful and prefer_reference and reference_available():
        from oracle import ref_py
        key = id(case["mesh"])
        if key not in _REF_CACHE:
            _REF_CACHE.clear()
            _REF_CACHE[key] = ref_py.Reference(case["mesh"])
        ref = _REF_CACHE[key]
        texts = [ref_py.dict_text_from_record(case["shapes"][int(sh)]) for sh in solids["shape"]]
        seeds = np.array([o.nearest_cell(solids[i]["pos"]) if i < n_sample else 0 for i in range(len(solids))], dtype=np.int32)
        for _ in range(repeats):
            r = ref.interact(texts, solids, seeds, case["U"], case["dt"], case["rhof"], case["two_d"], solid_range=(0, n_sample),
                             want_lists=False)
            if best is None or r["timing_ms"] < best[1]:
                best = (r["pairs"], r["timing_ms"], "reference")
        return best
    for _ in range(repeats):
        r = o.interact(case["shapes"], solids, case["U"], case["dt"], case["rhof"], faithful=faithful,
                       want_lists=True, solid_range=(0, n_sample))
        pairs = int(r["list_off"][3 * n_sample])
        ms = float(r["timing_ms"][0])
        if best is None or ms < best[1]:
            best = (pairs, ms, "port")
    return best


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU implementation of interact() (its compiled classes, oracle/_ref; the oracle
    port where that library is absent) on the host cores, a bounded sample of the workload per step."""
    if rank != 0:
        return
    wl, case, desc = build_case(args, 0, 1 if args.workload != "c5" and world == 1 else world)
    # Bounded sample: the per-solid cost of the faithful algorithm is O(nCells) (a fresh CELL_TYPE array per solid, a whole-mesh
    # scan for every solid that has no vertex-inside cell on this rank), i.e. seconds per solid on a C5 block.  Two solids are
    # probed; the run's time box (--cpu-budget seconds, default 150) then fixes how many solid evaluations fit.  At least
    # MIN_DISTINCT distinct solids are timed over the run: when a step cannot hold that many, every step takes a DIFFERENT
    # slice of one evenly spaced sample, and the line's value is total pairs / total time over the timed steps.
    MIN_DISTINCT = 32
    n_total = args.warmup + args.steps
    prefer = True
    try:
        _, probe_ms, kind = cpu_baseline_sample(case, 2, faithful=True, indices=representative_solids(case, 2))
    except Exception as ex:   # the compiled reference (oracle/_ref) is optional
        print(f"[bench] compiled reference unavailable ({ex}); timing the oracle port", file=sys.stderr)
        prefer = False
        _, probe_ms, kind = cpu_baseline_sample(case, 2, faithful=True, indices=representative_solids(case, 2), prefer_reference=False)
    per_solid_s = max(probe_ms * 1e-3 / 2.0, 1e-4)
    evals = max(1.0, args.cpu_budget / per_solid_s)                 # solid evaluations the time box holds
    per_step = int(max(1, min(args.cpu_solids, evals // max(n_total, 1))))
    rotate = per_step < MIN_DISTINCT
    pool = representative_solids(case, max(per_step, min(MIN_DISTINCT, len(case["solids"]))) if rotate else per_step)
    if rotate:   # a step cannot hold MIN_DISTINCT solids: warm-up = 1 solid, the timed steps share the pool between them
        per_step = max(1, int(np.ceil(len(pool) / max(args.steps, 1))))
        per_step = int(min(per_step, max(1, (evals - args.warmup) // max(args.steps, 1))))
    vals, used = [], set()
    for i in range(n_total):
        if not rotate:
            idx = pool
        elif i < args.warmup:
            idx = pool[:1]
        else:
            j = (i - args.warmup) * per_step
            idx = pool[[(j + t) % len(pool) for t in range(per_step)]]
        pairs, ms, kind = cpu_baseline_sample(case, len(idx), faithful=True, indices=idx, prefer_reference=prefer)
        if i >= args.warmup:
            vals.append((pairs, ms))
            used.update(int(x) for x in idx)
    pairs = sum(p for p, _ in vals)
    ms = sum(m for _, m in vals)
    value = pairs / (ms * 1e-3)
    n_distinct = len(used)
    sample = (f"{per_step} of {len(case['solids'])} solids per step"
              + (f", a different slice of one evenly spaced {len(pool)}-solid sample every step ({n_distinct} distinct solids over the {args.steps} timed steps)" if rotate else "")
              + f" (evenly spaced over the solids that touch rank 0's block and those that do not, in their proportion; sized by a 2-solid probe "
              f"of {per_solid_s:.2f} s per solid to a {args.cpu_budget:.0f} s run) on the full mesh of rank 0 of {world}, timed region = solid loop + "
              f"checkAlpha (reference src/solidcloud.cpp:442-451), "
              + ("the reference's own compiled CellEnumerator / GeometricTools / Solid / shape classes (oracle/_ref) inside the loop of "
                 "src/solidcloud.cpp:361-464; " if kind == "reference" else "oracle port, faithful per-solid O(nCells) CELL_TYPE array; ")
              + f"1 process, 1 thread "
              f"(the reference is single-threaded per MPI rank; see DESIGN.md for why an MPI split is slower on this path)")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / max(1, len(vals)), "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": make_config(desc),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": kind, "sample": sample, "distinct_solids": n_distinct},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    try:   # the product's CUDA library takes no part in this arm: only the mesh helper (libsdfibm_mesh.so) and the checker are mapped
        line["product_cuda_library_mapped"] = "libsdfibm_b200.so" in open("/proc/self/maps").read()
    except OSError:
        pass
    if n_distinct < MIN_DISTINCT and n_distinct < len(case["solids"]):
        line["ratio"] = None
        line["ratio_reason"] = (f"only {n_distinct} distinct solids fit the {args.cpu_budget:.0f} s time box at {per_solid_s:.1f} s per solid: "
                                f"too few for a meaningful updates/s figure; use the N=1 line's ratio")
    print(json.dumps(line), flush=True)


def parity_check(case, ctx, got, partial_ft, n_check, use_reference=True):
    """What the timed kernels produced (fields and per-solid sums of the last step, host copies) against the CPU checker on
    `n_check` evenly spaced solids that touch this rank's block: the reference's own compiled classes (oracle/_ref) when
    available and the mesh is small enough for their O(nCells)-per-solid cost, else the oracle port (same results).
    Candidate lists must be equal; As / Fs are compared on the sample's cells that no other solid touches; force / torque
    relative to the largest component of the sample."""
    from oracle.oracle_py import Oracle

    S = case["solids"]
    idx = representative_solids(case, n_check)
    rb = np.array([float(sh["radiusB"]) for sh in case["shapes"]])
    m = case["mesh"]
    r = np.where(rb[S["shape"][idx]] > 0, rb[S["shape"][idx]], 0.0) + 1.0
    idx = idx[np.all((S["pos"][idx] + r[:, None] >= m.bounds_min) & (S["pos"][idx] - r[:, None] <= m.bounds_max), axis=1)]
    sub = np.ascontiguousarray(S[idx])
    o = Oracle(m, case["two_d"])
    kind = "port"
    if use_reference and reference_available() and m.n_cells <= 20_000_000:
        from oracle import ref_py
        key = id(m)
        if key not in _REF_CACHE:
            _REF_CACHE.clear()
            _REF_CACHE[key] = ref_py.Reference(m)
        texts = [ref_py.dict_text_from_record(case["shapes"][int(sh)]) for sh in sub["shape"]]
        seeds = np.array([o.nearest_cell(x) for x in sub["pos"]], dtype=np.int32)
        ref = _REF_CACHE[key].interact(texts, sub, seeds, case["U"], case["dt"], case["rhof"], case["two_d"])
        kind = "reference"
    else:
        ref = o.interact(case["shapes"], sub, case["U"], case["dt"], case["rhof"])
    off, cells = ctx.candidate_lists()
    lists_equal = True
    for j, s_ in enumerate(idx):
        for t in range(3):
            a = cells[off[3 * s_ + t]: off[3 * s_ + t + 1]]
            b = ref["list_cells"][ref["list_off"][3 * j + t]: ref["list_off"][3 * j + t + 1]]
            lists_equal = lists_equal and np.array_equal(a, b)
    owners = np.bincount(cells, minlength=m.n_cells)
    mine = np.unique(ref["list_cells"])
    mine = mine[owners[mine] == 1]
    def rel(a, b):
        d = np.abs(a - b)
        nz = np.abs(b) > 0
        return float(np.max(d[nz] / np.abs(b[nz]))) if nz.any() else 0.0
    # the checker numbered the sampled solids 0..k-1: ALL_INSIDE cells carry (solid id + 4) (solidcloud.cpp:376-382)
    ct_ref = ref["Ct"].copy()
    ai = ct_ref >= 4
    ct_ref[ai] = idx[(ct_ref[ai] - 4).astype(np.int64)] + 4.0
    ft_scale = max(float(np.abs(ref["FT"]).max()), 1e-300)
    return {"checker": kind, "solids_checked": int(len(idx)), "pairs_checked": int(ref["list_off"][-1]), "lists_equal": bool(lists_equal),
            "single_owner_cells": int(len(mine)),
            "max_rel_As": rel(got["As"][mine], ref["As"][mine]), "max_rel_Ts": rel(got["Ts"][mine], ref["Ts"][mine]),
            "max_rel_Fs": float(np.abs(got["Fs"][mine] - ref["Fs"][mine]).max() / max(np.abs(ref["Fs"][mine]).max(), 1e-300)) if len(mine) else 0.0,
            "Ct_equal": bool(np.array_equal(got["Ct"][mine], ct_ref[mine])),
            "max_rel_FT": float(np.abs(partial_ft[idx] - ref["FT"]).max() / ft_scale),
            "FT_note": "this rank's partial sums (before the all-reduce) vs the checker on this rank's block; relative to the sample's largest component"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="auto", choices=["auto", "c1", "c2", "c3", "c3b", "c3tc", "c4", "c5"],
                    help="auto = the contract's line (C4 at N=1, C5 at N>1); c1..c3b = the reference's small example set-ups (ms/step)")
    ap.add_argument("--cells-per-side", dest="n", type=int, default=0, help="cells per side (scaled-down runs)")
    ap.add_argument("--solids", type=int, default=0)
    ap.add_argument("--cpu-solids", type=int, default=256, help="solids per CPU-baseline sample")
    ap.add_argument("--cpu-budget", type=float, default=150.0, help="--impl reference: seconds of CPU work for the whole run")
    ap.add_argument("--no-base", action="store_true", help="N>1: skip the same-workload 1-GPU base run (scaling_base)")
    ap.add_argument("--no-check", action="store_true", help="skip the in-bench parity check against the CPU checker")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist

    import __graft_entry__ as ge
    ge.build()
    from sdfibm_b200.context import Context

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: there is no CPU fallback for the product path")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    wl, case, desc = build_case(args, rank, world)
    mesh = case["mesh"]
    nC, nS = mesh.n_cells, len(case["solids"])
    ctx = Context(local_rank)
    ctx.set_mesh(mesh, case["two_d"])
    ctx.set_shapes(case["shapes"])
    dev = torch.device("cuda", local_rank)
    ext = torch.cuda.ExternalStream(ctx.stream_ptr(), device=dev)

    dU = torch.from_numpy(case["U"]).to(dev)
    dAs = torch.empty(nC, dtype=torch.float64, device=dev)
    dFs = torch.empty(nC, 3, dtype=torch.float64, device=dev)
    dTs = torch.empty(nC, dtype=torch.float64, device=dev)
    dCt = torch.empty(nC, dtype=torch.float64, device=dev)
    dFT = torch.empty(nS, 6, dtype=torch.float64, device=dev)
    torch.cuda.synchronize()

    from sdfibm_b200 import capi
    solids_pinned = capi.pinned_like(np.ascontiguousarray(case["solids"], dtype=capi.SOLID_DTYPE))   # the host side's own solid array

    if world > 1:
        # the library's own NCCL communicator (C ABI: sdfibm_comm_init): from here on every interact uploads this rank's 1/N slice
        # of the replicated solid array, all-gathers the slices over NVLink, and all-reduces the per-solid force/torque on the
        # context stream right behind the kernels (replaces the 2N Foam::reduce calls of solidcloud.cpp:427-431)
        from sdfibm_b200 import parallel
        parallel.init_library_comm(ctx)

    comm_ms = []

    def step_device():
        ctx.interact_device(solids_pinned, dU.data_ptr(), case["dt"], case["rhof"], dAs.data_ptr(), dFs.data_ptr(),
                            dTs.data_ptr(), dCt.data_ptr(), dFT.data_ptr())
        if world > 1:
            comm_ms.append(ctx.comm_last_ms())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    split = []
    host_us = []

    def timed(fn, steps, warmup):
        with torch.cuda.stream(ext):
            for _ in range(warmup):
                fn()
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            kern_ms, pipe_ms = [], []
            split.clear()
            host_us.clear()
            e0.record(ext)
            for _ in range(steps):
                fn()                 # nothing but the step between the two events: a C host has no interpreter gaps either
            e1.record(ext)
            barrier()
            ms = e0.elapsed_time(e1)
            # the per-kernel split (the library's own CUDA events inside the graph launch) comes from further, untimed steps
            for _ in range(min(steps, 5)):
                fn()
                t = ctx.last_timings()
                kern_ms.append(t["interact_kernels_ms"])
                split.append((t["classify_ms"], t["heavy_ms"], t["final_ms"], t["connectivity_ms"], t["binning_ms"]))
                pipe_ms.append(t["pipeline_ms"])
                host_us.append(list(ctx.last_host_timings().values()))
            barrier()
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, float(np.mean(kern_ms)), float(np.mean(pipe_ms))

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms, kern_ms, pipe_ms = timed(step_device, args.steps, args.warmup)
    clocks = sampler.stop() if rank == 0 else None
    split_mean = [float(x) for x in np.mean(np.array(split), axis=0)]
    host_mean = [float(x) for x in np.mean(np.array(host_us), axis=0)]   # of the device-resident steps (the e2e loop below reuses the list)
    counts = ctx.candidate_counts()
    stats = ctx.last_stats()
    launches_per_step = stats["launches"]
    tot = torch.tensor(counts + [nC], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tot)
    counts_all = [int(x) for x in tot[:3].tolist()]
    pairs_all = sum(counts_all)
    ms_per_step = ms / args.steps
    value = pairs_all / (ms_per_step * 1e-3)

    # ---- end to end through the host-buffer C-ABI entry (pinned host arrays) ----
    e2e = None
    if not args.no_e2e:
        hU = torch.from_numpy(case["U"]).pin_memory()
        hAs = torch.empty(nC, dtype=torch.float64).pin_memory()
        hFs = torch.empty(nC, 3, dtype=torch.float64).pin_memory()
        hTs = torch.empty(nC, dtype=torch.float64).pin_memory()
        hCt = torch.empty(nC, dtype=torch.float64).pin_memory()
        hFT = torch.empty(nS, 6, dtype=torch.float64).pin_memory()
        out = dict(As=hAs.numpy(), Fs=hFs.numpy(), Ts=hTs.numpy(), Ct=hCt.numpy(), FT=hFT.numpy())
        hUn = hU.numpy()

        def step_host():
            # (N > 1: force/torque comes back summed over the ranks — the library's all-reduce, solidcloud.cpp:427-431)
            ctx.interact(solids_pinned, hUn, case["dt"], case["rhof"], out=out)

        e_steps = max(2, min(args.steps, 5))
        ms_e, _, _ = timed(step_host, e_steps, max(1, min(args.warmup, 3)))
        e_ms_step = ms_e / e_steps
        h2d = int(hUn.nbytes + case["solids"].nbytes)
        d2h = int(sum(v.nbytes for v in out.values()))
        e2e = {"value": pairs_all / (e_ms_step * 1e-3), "unit": UNIT, "ms_per_step": e_ms_step,
               "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": e_steps,
               "note": "sdfibm_interact (host-buffer C ABI): pinned U H2D, kernels, As/Fs/Ts/Ct/forceTorque D2H every step; "
                       "the copies stream in cell chunks on two copy streams, overlapped with the kernels and with each other"}

    # ---- end to end with the flow fields resident in HBM: per step the host sends the solid states and reads the
    #      per-solid force/torque back; interact and the forcing it drives (U -= Fs dt, T = (1 - As) T + Ts, main.cpp:70-77) run on
    #      the device, As / Fs / Ts / Ct stay there for the caller's flux / pressure terms ----
    e2e_res = None
    touched = None
    if not args.no_e2e:
        dUr = dU.clone()
        dT = torch.full((nC,), 300.0, dtype=torch.float64, device=dev)
        hFTr_np = capi.pinned_like(np.zeros((nS, 6)))

        def step_resident():
            ctx.interact_device(solids_pinned, dUr.data_ptr(), case["dt"], case["rhof"], dAs.data_ptr(), dFs.data_ptr(),
                                dTs.data_ptr(), dCt.data_ptr(), dFT.data_ptr())          # H2D: the solid records (pinned)
            ctx.apply_forcing_device(dUr.data_ptr(), dT.data_ptr(), case["dt"])
            ctx.download(hFTr_np, dFT.data_ptr())                                        # D2H: force / torque; the step's one host sync

        r_steps = max(3, min(args.steps, 10))
        ms_r, _, _ = timed(step_resident, r_steps, max(1, min(args.warmup, 3)))
        r_ms_step = ms_r / r_steps
        e2e_res = {"value": pairs_all / (r_ms_step * 1e-3), "unit": UNIT, "ms_per_step": r_ms_step,
                   "h2d_bytes_per_step": int(case["solids"].nbytes), "d2h_bytes_per_step": int(hFTr_np.nbytes + 80), "steps": r_steps,
                   "note": "fields resident in HBM (row f2): sdfibm_interact_device (pinned solid records H2D inside) + sdfibm_apply_forcing_device "
                           "(U -= Fs dt, T = (1 - As) T + Ts on the device) + force/torque and the 80-byte status word D2H, one host "
                           "synchronisation per step; As / Fs / Ts / Ct stay on the device"}
        if rank == 0:
            t0 = time.perf_counter()
            tc = ctx.touched_cells()
            t1 = time.perf_counter()
            touched = {"cells": int(len(tc["cells"])), "bytes": int(sum(v.nbytes for v in tc.values())), "ms": (t1 - t0) * 1e3,
                       "note": "sdfibm_touched_cells: compact (cell, As, Fs, Ts, Ct) records of the touched cells to pageable host arrays, "
                               "one call, wall clock (for hosts that keep their own copy of the fields)"}
            del tc
        del dUr, dT
        # restore the state the parity check below expects (the resident steps changed U)
        ctx.interact_device(solids_pinned, dU.data_ptr(), case["dt"], case["rhof"], dAs.data_ptr(), dFs.data_ptr(),
                            dTs.data_ptr(), dCt.data_ptr(), dFT.data_ptr())

    # ---- what the timed kernels produced, checked against the CPU checker (rank 0's block) ----
    check = None
    if world > 1:
        # one extra untimed step WITHOUT the all-reduce: this rank's partial sums; their NCCL sum must reproduce the timed step's dFT
        total_timed = dFT.clone()
        ctx.comm_options(auto_reduce=False, gather_solids=True)
        ctx.interact_device(solids_pinned, dU.data_ptr(), case["dt"], case["rhof"], dAs.data_ptr(), dFs.data_ptr(),
                            dTs.data_ptr(), dCt.data_ptr(), dFT.data_ptr())
        ctx.comm_options(auto_reduce=True, gather_solids=True)
        partial = dFT.clone()
        total_again = partial.clone()
        dist.all_reduce(total_again)   # torch's own NCCL as the independent check of the library's all-reduce
        allreduce_dev = float((total_again - total_timed).abs().max().item() / max(float(total_timed.abs().max().item()), 1e-300))
    else:
        partial = dFT
        allreduce_dev = None
    if rank == 0 and not args.no_check:
        got = {"As": dAs.cpu().numpy(), "Fs": dFs.cpu().numpy(), "Ts": dTs.cpu().numpy(), "Ct": dCt.cpu().numpy()}
        try:
            check = parity_check(case, ctx, got, partial.cpu().numpy(), 64 if world == 1 else 32)
            check["counts_equal_device_and_lists"] = bool(sum(counts) == int(ctx.candidate_lists()[0][-1]))
            if allreduce_dev is not None:
                check["allreduce_vs_sum_of_partials_rel"] = allreduce_dev
        except Exception as ex:   # the checker must never take the bench line down
            check = {"error": repr(ex)}
        del got

    line = None
    if rank == 0:
        peak, peak_src = measured_hbm_peak()
        alg = algorithmic_bytes(nC, counts, nS)            # rank 0's kernel launch
        # one launch = one CUDA-graph launch of the whole interact pipeline (binning, k_classify, k_heavy, k_final, k_connectivity,
        # k_finalize), timed live by the library's CUDA events on its stream
        achieved = alg / (pipe_ms * 1e-3) / 1e9
        traffic, traffic_src = measured_traffic(wl if world == 1 and not args.n else "-")
        heavy_share = split_mean[1] / max(pipe_ms, 1e-30)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": make_config(desc),
            "workload_stats": dict(cells_total=int(tot[3].item()), solids=nS, pairs_per_step=pairs_all,
                                   pairs_by_type={"ALL_INSIDE": counts_all[0], "CENTER_INSIDE": counts_all[1], "CENTER_OUTSIDE": counts_all[2]}),
            "clocks": clocks,
            "gpu_launches": int(launches_per_step * args.steps),
            "kernel_ms": {"k_classify": split_mean[0], "k_heavy": split_mean[1], "k_final": split_mean[2],
                          "k_connectivity+finalise": split_mean[3], "solid_binning": split_mean[4],
                          "interact_kernels": kern_ms, "pipeline_device": pipe_ms, "step_wall_on_stream": ms_per_step,
                          "heavy_items": stats["heavy_items"],
                          "comm_device_ms": (float(np.mean(comm_ms[-args.steps:])) if comm_ms else None),
                          "host_us": dict(zip(("stage_solids", "enqueue", "wait_gpu", "call"), host_mean))},
            "roofline": {"bound": "hbm", "kernel": f"one graph launch of the interact pipeline: binning + k_classify + k_heavy + k_final + k_connectivity (k_heavy is {100 * heavy_share:.0f}% of it)", "achieved": achieved,
                         "achieved_interact_kernels_only": alg / (kern_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic,
                         "traffic_source": traffic_src["source"] if traffic_src else None,
                         "traffic_per_kernel": traffic_src["kernels"] if traffic_src else None, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": int(alg),
                         "formula": "48*nCells + 112*P + 216*P_b + 176*N (SURVEY.md 8d), rank 0"},
            "e2e": e2e_res,
            "e2e_host_fields": e2e,
            "touched_download": touched,
            "parity_check": check,
        }
        if not args.no_cpu and world == 1:
            t0 = time.time()
            try:
                pairs_s, ms_s, kind = cpu_baseline_sample(case, args.cpu_solids, faithful=True)
            except Exception as ex:   # the compiled reference (oracle/_ref) is optional: fall back to the oracle port
                print(f"[bench] compiled reference unavailable ({ex}); timing the oracle port", file=sys.stderr)
                pairs_s, ms_s, kind = cpu_baseline_sample(case, args.cpu_solids, faithful=True, prefer_reference=False)
            pairs_f, ms_f, _ = cpu_baseline_sample(case, args.cpu_solids, faithful=False)
            what = ("the reference's own compiled CellEnumerator / GeometricTools / Solid / shape classes (oracle/_ref) inside the loop of "
                    "src/solidcloud.cpp:361-464" if kind == "reference" else "oracle port")
            line["cpu_baseline"] = {
                "value": pairs_s / (ms_s * 1e-3), "unit": UNIT, "cores": 1, "kind": kind,
                "sample": f"first {args.cpu_solids} of {nS} solids, one step, {what}; region = solid loop + checkAlpha "
                          f"(reference src/solidcloud.cpp:442-451), faithful per-solid O(nCells) CELL_TYPE array; "
                          f"{pairs_s} pairs in {ms_s:.1f} ms; host has {os.cpu_count()} cores, reference uses 1 per rank",
                "optimised_port_value": pairs_f / (ms_f * 1e-3),
                "optimised_port_note": "same oracle without the per-solid O(nCells) allocation (not what the reference does)",
                "wall_s": time.time() - t0,
            }
            line["ratios_vs_optimised_port"] = {
                "device_resident": value / line["cpu_baseline"]["optimised_port_value"],
                "e2e": (e2e_res["value"] / line["cpu_baseline"]["optimised_port_value"]) if e2e_res else None,
                "e2e_host_fields": (e2e["value"] / line["cpu_baseline"]["optimised_port_value"]) if e2e else None,
                "note": "the faithful CPU path pays a 67 MB CELL_TYPE allocation per solid (src/cellenumerator.cpp:49); these are the ratios against "
                        "the same algorithm without it, to be quoted beside the driver's headline ratio"}
    # orderly teardown: our context (own CUDA stream / events) before torch's NCCL communicator and allocator
    torch.cuda.synchronize()
    ctx.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()

    # ---- N > 1: the SAME workload on ONE GPU, in the same run (rank 0, after the other ranks have left): the base of the scaling figure
    if rank == 0 and world > 1 and not args.no_base:
        try:
            del dU, dAs, dFs, dTs, dCt, dFT, partial, total_timed, total_again, case, mesh
            import gc
            gc.collect()
            torch.cuda.empty_cache()
            if wl != "c5":
                raise RuntimeError("no single-GPU base for this workload")
            case1, desc1 = _c5_single(args)
            base = device_resident_base(case1, local_rank, steps=max(2, min(args.steps, 5)), warmup=3)
            line["scaling_base"] = {"workload": desc1["workload"], "n_gpus": 1, "ms_per_step": base["ms_per_step"],
                                    "pairs_per_step": base["pairs"], "value": base["value"],
                                    "note": "the SAME mesh and solids on one GPU of this box, timed in this run by rank 0 after the N-rank measurement"}
            line["speedup_same_workload"] = base["ms_per_step"] / ms_per_step
            line["pairs_match_base"] = bool(base["pairs"] == pairs_all)
        except Exception as ex:
            line["scaling_base"] = {"error": repr(ex)}
    if rank == 0:
        print(json.dumps(line), flush=True)
    del ext
    sys.stdout.flush()


def _c5_single(args):
    """The whole C5 mesh as ONE block (the 1-GPU base of the N>1 lines)."""
    from sdfibm_b200 import cases
    n = args.n or 512
    scale = n / 512.0
    n_side = max(1, int(round(47 * scale)))
    n_solids = args.solids or min(n_side ** 3, int(round(100000 * scale ** 3)))
    case = cases.case_c5_block(0, 1, n=n, n_solids=n_solids, n_side=n_side)
    desc = {"workload": f"C5: {n}^3 hex cells, {n_solids} Sphere r=4.5 / Ellipsoid (5,4.5,4) mixed, jittered {n_side}^3 lattice, seed 12345; one block"}
    return case, desc


def device_resident_base(case, local_rank, steps, warmup):
    """Device-resident ms/step of one case on one GPU (CUDA events on the library's stream)."""
    import torch
    from sdfibm_b200 import capi
    from sdfibm_b200.context import Context

    dev = torch.device("cuda", local_rank)
    mesh = case["mesh"]
    nC, nS = mesh.n_cells, len(case["solids"])
    ctx = Context(local_rank)
    ctx.set_mesh(mesh, case["two_d"])
    ctx.set_shapes(case["shapes"])
    ext = torch.cuda.ExternalStream(ctx.stream_ptr(), device=dev)
    dU = torch.from_numpy(case["U"]).to(dev)
    f = [torch.empty(n, dtype=torch.float64, device=dev) for n in (nC, 3 * nC, nC, nC, 6 * nS)]
    solids = capi.pinned_like(np.ascontiguousarray(case["solids"], dtype=capi.SOLID_DTYPE))
    torch.cuda.synchronize()

    def step():
        ctx.interact_device(solids, dU.data_ptr(), case["dt"], case["rhof"], *[x.data_ptr() for x in f])

    with torch.cuda.stream(ext):
        for _ in range(warmup):
            step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(ext)
        for _ in range(steps):
            step()
        e1.record(ext)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
    pairs = sum(ctx.candidate_counts())
    ctx.close()
    del ext
    return {"ms_per_step": ms, "pairs": int(pairs), "value": pairs / (ms * 1e-3)}


if __name__ == "__main__":
    main()
