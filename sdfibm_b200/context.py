"""Thin Python host over the C ABI (one Context per rank/GPU).

Mirrors the per-step surface of sdfibm::SolidCloud that main.cpp uses (reference src/solidcloud.h:98-116):
``interact`` / ``fix_internal`` plus the collision step, on numpy host arrays or raw device pointers.
All compute happens in libsdfibm_b200.so; nothing here falls back to the CPU.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import capi


class Context:
    def __init__(self, device: int = 0, cell_slots: int | None = None, allow_order_free: bool = False):
        self._lib = capi.load()
        self._h = C.c_void_p()
        capi.check(self._lib.sdfibm_create(int(device), C.byref(self._h)))
        if cell_slots is not None:
            capi.check(self._lib.sdfibm_set_cell_slots(self._h, int(cell_slots)))
        if allow_order_free:   # meshes whose cells differ in vertex count: accept the order-free ALL_INSIDE rule
            capi.check(self._lib.sdfibm_allow_order_free(self._h, 1))
        self.mesh = None
        self.n_cells = 0
        self.n_solids = 0

    def close(self):
        if self._h:
            self._lib.sdfibm_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- one-time uploads ----
    def set_mesh(self, mesh, two_d: bool):
        capi.check(self._lib.sdfibm_set_mesh(self._h, C.byref(mesh.view), int(bool(two_d))))
        self.mesh = mesh
        self.n_cells = mesh.n_cells

    def set_shape_programs(self, ops: np.ndarray):
        """Op table of the composed shapes (SDFIBM_SHAPE_PROGRAM records point into it); call before set_shapes."""
        ops = np.ascontiguousarray(ops, dtype=capi.SDF_OP_DTYPE)
        capi.check(self._lib.sdfibm_set_shape_programs(self._h, capi.ptr(ops), len(ops)))

    def set_shapes(self, shapes: np.ndarray):
        shapes = np.ascontiguousarray(shapes, dtype=capi.SHAPE_DTYPE)
        capi.check(self._lib.sdfibm_set_shapes(self._h, capi.ptr(shapes), len(shapes)))

    # ---- SolidCloud::interact ----
    def interact(self, solids: np.ndarray, U: np.ndarray, dt: float, rhof: float, out=None):
        """Host-buffer interact.  Returns dict(As, Fs, Ts, Ct, FT)."""
        solids = np.ascontiguousarray(solids, dtype=capi.SOLID_DTYPE)
        U = np.ascontiguousarray(U, dtype=np.float64)
        nC, n = self.n_cells, len(solids)
        assert U.size == 3 * nC
        if out is None:
            out = dict(As=np.empty(nC), Fs=np.empty((nC, 3)), Ts=np.empty(nC), Ct=np.empty(nC), FT=np.empty((max(n, 0), 6)))
        capi.check(self._lib.sdfibm_interact(self._h, capi.ptr(solids), n, capi.ptr(U), float(dt), float(rhof),
                                             capi.ptr(out["As"]), capi.ptr(out["Fs"]), capi.ptr(out["Ts"]),
                                             capi.ptr(out["Ct"]), capi.ptr(out["FT"])))
        self.n_solids = n
        return out

    def interact_device(self, solids: np.ndarray, dU: int, dt: float, rhof: float, dAs: int, dFs: int, dTs: int,
                        dCt: int, dFT: int):
        """Device-pointer interact (pointers as ints, e.g. torch.Tensor.data_ptr())."""
        solids = np.ascontiguousarray(solids, dtype=capi.SOLID_DTYPE)
        capi.check(self._lib.sdfibm_interact_device(self._h, capi.ptr(solids), len(solids), capi.ptr(dU), float(dt),
                                                    float(rhof), capi.ptr(dAs), capi.ptr(dFs), capi.ptr(dTs),
                                                    capi.ptr(dCt), capi.ptr(dFT)))
        self.n_solids = len(solids)

    def interact_device_solids(self, d_solids: int, n_solids: int, dU: int, dt: float, rhof: float, dAs: int, dFs: int, dTs: int,
                               dCt: int, dFT: int, may_be_global: bool = True):
        """Device-pointer interact with the solid records already on the device (pointer as int)."""
        capi.check(self._lib.sdfibm_interact_device_solids(self._h, capi.ptr(d_solids), int(n_solids), int(bool(may_be_global)),
                                                           capi.ptr(dU), float(dt), float(rhof), capi.ptr(dAs), capi.ptr(dFs),
                                                           capi.ptr(dTs), capi.ptr(dCt), capi.ptr(dFT)))
        self.n_solids = int(n_solids)

    # ---- the step either side of interact, on the device (main.cpp:70-77) ----
    def apply_forcing_device(self, dU: int | None, dT: int | None, dt: float):
        """U -= Fs dt, T = (1 - As) T + Ts on the device arrays (pointers as ints; None skips one), from the last interact_device."""
        capi.check(self._lib.sdfibm_apply_forcing_device(self._h, capi.ptr(dU) if dU else None, capi.ptr(dT) if dT else None, float(dt)))

    def download(self, host: np.ndarray, device_ptr: int):
        """Stream-ordered device -> host copy into `host` (ideally page-locked), then wait: the step's one synchronisation."""
        capi.check(self._lib.sdfibm_download(self._h, capi.ptr(host), capi.ptr(device_ptr), host.nbytes))

    def touched_cells(self):
        """Compact host records of the cells the last interact touched: dict(cells, As, Fs, Ts, Ct)."""
        n = C.c_int64()
        capi.check(self._lib.sdfibm_touched_cells(self._h, 0, C.byref(n), None, None, None, None, None))
        k = int(n.value)
        out = dict(cells=np.empty(k, np.int32), As=np.empty(k), Fs=np.empty((k, 3)), Ts=np.empty(k), Ct=np.empty(k))
        if k:
            capi.check(self._lib.sdfibm_touched_cells(self._h, k, C.byref(n), capi.ptr(out["cells"]), capi.ptr(out["As"]), capi.ptr(out["Fs"]),
                                                      capi.ptr(out["Ts"]), capi.ptr(out["Ct"])))
        return out

    # ---- cross-rank exchange (NCCL inside the library) ----
    @staticmethod
    def comm_unique_id() -> bytes:
        """rank 0: the 128-byte NCCL id every rank passes to comm_init (broadcast it with whatever the host has)."""
        buf = C.create_string_buffer(128)
        capi.check(capi.load().sdfibm_comm_unique_id(buf))
        return buf.raw

    def comm_init(self, uid: bytes, rank: int, n_ranks: int):
        """Collective.  From here on interact / interact_device return force/torque summed over the ranks (one ncclAllReduce on
        the context stream, src/solidcloud.cpp:427-431) and upload only this rank's slice of the replicated solid array."""
        assert len(uid) == 128
        capi.check(self._lib.sdfibm_comm_init(self._h, C.c_char_p(uid), int(rank), int(n_ranks)))

    def comm_options(self, auto_reduce: bool = True, gather_solids: bool = True):
        capi.check(self._lib.sdfibm_comm_options(self._h, int(bool(auto_reduce)), int(bool(gather_solids))))

    def comm_destroy(self):
        capi.check(self._lib.sdfibm_comm_destroy(self._h))

    def allreduce_force_torque(self, dFT: int, n_solids: int):
        """In-place sum over the ranks of a device array [6 n_solids], stream-ordered on the context stream."""
        capi.check(self._lib.sdfibm_allreduce_force_torque(self._h, capi.ptr(dFT), int(n_solids)))

    def comm_last_ms(self) -> float:
        ms = C.c_double()
        capi.check(self._lib.sdfibm_comm_last_ms(self._h, C.byref(ms)))
        return float(ms.value)

    # ---- SolidCloud::fixInternal ----
    def fix_internal(self, solids: np.ndarray, U: np.ndarray) -> np.ndarray:
        solids = np.ascontiguousarray(solids, dtype=capi.SOLID_DTYPE)
        U = np.array(U, dtype=np.float64, copy=True, order="C")
        capi.check(self._lib.sdfibm_fix_internal(self._h, capi.ptr(solids), len(solids), capi.ptr(U)))
        return U

    def fix_internal_device(self, solids: np.ndarray, dU: int, dCt: int | None = None):
        solids = np.ascontiguousarray(solids, dtype=capi.SOLID_DTYPE)
        capi.check(self._lib.sdfibm_fix_internal_device(self._h, capi.ptr(solids), len(solids), capi.ptr(dU),
                                                        capi.ptr(dCt) if dCt else None))

    # ---- tool_vof: SolidCloud::writeVOF ----
    def volume_fraction(self, solids: np.ndarray):
        """(alpha[n_cells], sum(alpha V)) of the union of the solids (tool_vof/solidcloud.cpp:116-173)."""
        solids = np.ascontiguousarray(solids, dtype=capi.SOLID_DTYPE)
        alpha = np.empty(self.n_cells)
        tot = C.c_double()
        capi.check(self._lib.sdfibm_volume_fraction(self._h, capi.ptr(solids), len(solids), capi.ptr(alpha), C.byref(tot)))
        return alpha, float(tot.value)

    # ---- SolidCloud::calcMeanField ----
    def mean_field(self, solids: np.ndarray, field: np.ndarray):
        """Volume-weighted mean of `field` over each solid's (substitute) shape; returns (mean[N,3], sum_alpha_V[N])."""
        solids = np.ascontiguousarray(solids, dtype=capi.SOLID_DTYPE)
        field = np.ascontiguousarray(field, dtype=np.float64)
        n = len(solids)
        mean, den = np.empty((n, 3)), np.empty(n)
        capi.check(self._lib.sdfibm_mean_field(self._h, capi.ptr(solids), n, capi.ptr(field), capi.ptr(mean), capi.ptr(den)))
        return mean, den

    def mean_field_sums(self, solids: np.ndarray, field: np.ndarray):
        """The sampler's raw sums on this rank: (sum(alpha V field)[N,3], sum(alpha V)[N]) — reduce across ranks, then divide."""
        solids = np.ascontiguousarray(solids, dtype=capi.SOLID_DTYPE)
        field = np.ascontiguousarray(field, dtype=np.float64)
        n = len(solids)
        num, den = np.empty((n, 3)), np.empty(n)
        capi.check(self._lib.sdfibm_mean_field_sums(self._h, capi.ptr(solids), n, capi.ptr(field), capi.ptr(num), capi.ptr(den)))
        return num, den

    # ---- candidate lists / diagnostics of the last interact ----
    def candidate_counts(self):
        c = (C.c_int64 * 3)()
        capi.check(self._lib.sdfibm_candidate_counts(self._h, c))
        return [int(x) for x in c]

    def candidate_lists(self):
        n = self.n_solids
        off = np.zeros(3 * n + 1, dtype=np.int32)
        capi.check(self._lib.sdfibm_candidate_lists(self._h, capi.ptr(off), None, 0))
        cells = np.empty(max(int(off[-1]), 1), dtype=np.int32)
        capi.check(self._lib.sdfibm_candidate_lists(self._h, capi.ptr(off), capi.ptr(cells), len(cells)))
        return off, cells[: off[-1]]

    def last_stats(self):
        s = (C.c_int64 * 4)()
        capi.check(self._lib.sdfibm_last_stats(self._h, s))
        return dict(flagged_solids=int(s[0]), launches=int(s[1]), bin_entries=int(s[2]), heavy_items=int(s[3]))

    def last_host_timings(self):
        t = (C.c_double * 4)()
        capi.check(self._lib.sdfibm_last_host_timings(self._h, t))
        return {"stage_us": t[0], "enqueue_us": t[1], "wait_us": t[2], "call_us": t[3]}

    def last_aux_timings(self):
        t = (C.c_double * 2)()
        capi.check(self._lib.sdfibm_last_aux_timings(self._h, t))
        return {"fix_internal_ms": t[0], "collide_ms": t[1]}

    def last_timings(self):
        t = (C.c_double * 6)()
        capi.check(self._lib.sdfibm_last_timings(self._h, t))
        return dict(binning_ms=t[0], classify_ms=t[1], heavy_ms=t[2], final_ms=t[3], connectivity_ms=t[4],
                    pipeline_ms=t[5], interact_kernels_ms=t[1] + t[2] + t[3])

    def stream_ptr(self) -> int:
        p = C.c_void_p()
        capi.check(self._lib.sdfibm_stream(self._h, C.byref(p)))
        return int(p.value or 0)

    # ---- collision step ----
    def collide(self, solids: np.ndarray, delta: float, force_torque: np.ndarray | None = None, capacity: int = 1 << 16):
        solids = np.ascontiguousarray(solids, dtype=capi.SOLID_DTYPE)
        n = len(solids)
        ft = np.zeros((n, 6)) if force_torque is None else np.array(force_torque, dtype=np.float64, order="C")
        npairs = C.c_int64(0)
        cap = int(capacity)
        while True:
            pairs = np.empty((cap, 2), dtype=np.int32)
            ft_try = ft.copy()
            rc = self._lib.sdfibm_collide(self._h, capi.ptr(solids), n, float(delta), capi.ptr(pairs), cap,
                                          C.byref(npairs), capi.ptr(ft_try))
            if rc == 4 and npairs.value > cap:
                cap = int(npairs.value) + 16
                continue
            capi.check(rc)
            break
        return pairs[: npairs.value].copy(), ft_try

    def synchronize(self):
        capi.check(self._lib.sdfibm_synchronize(self._h))
