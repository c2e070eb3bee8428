"""Host-side mirror of the reference interface for the scheduling hot path.

`GpuScheduler.node_select(now, running, pending)` has the argument meaning and
error behaviour of `SchedulerAlgo::NodeSelect(now, running_jobs, pending_jobs)`
(reference: src/CraneCtld/JobScheduler.h:254-257): per-job failure is a pending
reason, never an exception; malformed input / missing device raises.

The work happens in the C-ABI library cranesched_b200/csrc/libcrane_sched.so
(hand-written sm_100a kernels). There is no CPU path: if the library is missing
or no CUDA device is present, construction fails loudly.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import abi

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libcrane_sched.so")
_libs: dict[str, C.CDLL] = {}


class CraneSchedError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"crane_sched error {code}: {msg}")
        self.code = code


def load_library(path: str | None = None) -> C.CDLL:
    path = path or os.environ.get("CRANE_SCHED_LIB") or LIB_PATH  # env override: A/B builds in tools/
    if path in _libs:
        return _libs[path]
    if not os.path.exists(path):
        raise FileNotFoundError(
            f"{path} not found: build it with `python -m cranesched_b200.build` "
            "(nvcc, sm_100a). There is no CPU fallback.")
    lib = C.CDLL(path)
    P = C.POINTER
    lib.crane_sched_create.restype = C.c_int
    lib.crane_sched_create.argtypes = [P(abi.SchedConfig), C.c_int, P(C.c_void_p)]
    lib.crane_sched_destroy.restype = None
    lib.crane_sched_destroy.argtypes = [C.c_void_p]
    lib.crane_sched_last_error.restype = C.c_char_p
    lib.crane_sched_last_error.argtypes = [C.c_void_p]
    lib.crane_sched_set_cluster.restype = C.c_int
    lib.crane_sched_set_cluster.argtypes = [C.c_void_p, P(abi.ClusterC)]
    lib.crane_sched_node_select.restype = C.c_int
    lib.crane_sched_node_select.argtypes = [C.c_void_p, C.c_int64, P(abi.RunningC), P(abi.PendingC), P(abi.PlacementsC)]
    lib.crane_sched_upload.restype = C.c_int
    lib.crane_sched_upload.argtypes = [C.c_void_p, P(abi.RunningC), P(abi.PendingC)]
    lib.crane_sched_run.restype = C.c_int
    lib.crane_sched_run.argtypes = [C.c_void_p, C.c_int64]
    lib.crane_sched_fetch.restype = C.c_int
    lib.crane_sched_fetch.argtypes = [C.c_void_p, P(abi.PlacementsC)]
    lib.crane_sched_sync.restype = C.c_int
    lib.crane_sched_sync.argtypes = [C.c_void_p, P(C.c_float)]
    lib.crane_sched_get_timing.restype = C.c_int
    lib.crane_sched_get_timing.argtypes = [C.c_void_p, P(abi.TimingC)]
    lib.crane_sched_qos_filter.restype = C.c_int
    lib.crane_sched_qos_filter.argtypes = [C.c_void_p, P(abi.QosTableC), C.c_void_p]
    lib.crane_sched_set_reservations.restype = C.c_int
    lib.crane_sched_set_reservations.argtypes = [C.c_void_p, P(abi.ReservationsC)]
    lib.crane_sched_pending_reset.restype = C.c_int
    lib.crane_sched_pending_reset.argtypes = [C.c_void_p]
    lib.crane_sched_pending_append.restype = C.c_int
    lib.crane_sched_pending_append.argtypes = [C.c_void_p, P(abi.PendingC), P(C.c_uint32)]
    lib.crane_sched_pending_erase.restype = C.c_int
    lib.crane_sched_pending_erase.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32]
    lib.crane_sched_set_running.restype = C.c_int
    lib.crane_sched_set_running.argtypes = [C.c_void_p, P(abi.RunningC)]
    lib.crane_sched_pending_rows.restype = C.c_uint32
    lib.crane_sched_pending_rows.argtypes = [C.c_void_p]
    lib.crane_sched_set_shard.restype = C.c_int
    lib.crane_sched_set_shard.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p]
    lib.crane_sched_device_placements.restype = C.c_int
    lib.crane_sched_device_placements.argtypes = [C.c_void_p, P(abi.DevicePlacementsC)]
    lib.crane_sched_debug_bitmap.restype = C.c_int
    lib.crane_sched_debug_bitmap.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, P(C.c_uint32), P(C.c_uint32)]
    lib.crane_sched_debug_profile.restype = C.c_int
    lib.crane_sched_debug_profile.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
    _libs[path] = lib
    return lib


EXPORTS = ("crane_sched_create", "crane_sched_destroy", "crane_sched_last_error",
           "crane_sched_set_cluster", "crane_sched_set_reservations", "crane_sched_node_select", "crane_sched_upload", "crane_sched_pending_reset", "crane_sched_pending_append", "crane_sched_pending_erase",
           "crane_sched_set_running", "crane_sched_pending_rows",
           "crane_sched_run", "crane_sched_fetch", "crane_sched_sync", "crane_sched_get_timing",
           "crane_sched_qos_filter", "crane_sched_set_shard", "crane_sched_device_placements",
           "crane_sched_debug_bitmap", "crane_sched_debug_profile")


class GpuScheduler:
    """One handle per GPU; owns the device-resident node, job and timeline tables."""

    def __init__(self, cfg: abi.Config, device: int = 0, lib_path: str | None = None):
        self._lib = load_library(lib_path)
        self._h = C.c_void_p()
        c_cfg = cfg.as_c()
        rc = self._lib.crane_sched_create(C.byref(c_cfg), device, C.byref(self._h))
        if rc != 0:
            raise CraneSchedError(rc, "crane_sched_create failed (no CUDA device?)" if rc == abi.ENODEV else "crane_sched_create")
        self.cfg = cfg
        self._keep = []

    def close(self):
        if self._h:
            self._lib.crane_sched_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int):
        if rc != 0:
            raise CraneSchedError(rc, (self._lib.crane_sched_last_error(self._h) or b"").decode())

    def set_cluster(self, cluster: abi.Cluster):
        c = cluster.as_c()
        self._check(self._lib.crane_sched_set_cluster(self._h, C.byref(c)))
        self.cluster = cluster

    def set_reservations(self, resv: "abi.Reservations | None"):
        """ResvMeta snapshot (JobScheduler.cpp:5655-5713); after set_cluster."""
        c = resv.as_c() if resv is not None else None
        self._check(self._lib.crane_sched_set_reservations(self._h, C.byref(c) if c is not None else None))

    # --- the NodeSelect call, host buffers in / host buffers out ------------
    def node_select(self, now: int, running: abi.Running, pending: abi.Pending,
                    out: abi.Placements | None = None) -> abi.Placements:
        out = out if out is not None else abi.Placements.for_pending(pending)
        c_rn, c_pd, c_out = running.as_c(), pending.as_c(), out.as_c()
        self._check(self._lib.crane_sched_node_select(self._h, now, C.byref(c_rn), C.byref(c_pd), C.byref(c_out)))
        return out

    # --- the same call split at the PCIe boundary ----------------------------
    def upload(self, running: abi.Running, pending: abi.Pending):
        c_rn, c_pd = running.as_c(), pending.as_c()
        self._keep = [running, pending]
        self._check(self._lib.crane_sched_upload(self._h, C.byref(c_rn), C.byref(c_pd)))

    # --- the pending table resident on the device across ticks ---------------------
    def pending_reset(self):
        self._check(self._lib.crane_sched_pending_reset(self._h))

    def pending_append(self, rows: abi.Pending) -> int:
        """Submit (JobScheduler.cpp:4254): returns the index of the first new row."""
        c, first = rows.as_c(), C.c_uint32(0)
        self._check(self._lib.crane_sched_pending_append(self._h, C.byref(c), C.byref(first)))
        return first.value

    def pending_erase(self, rows):
        r = np.ascontiguousarray(rows, np.uint32)
        self._check(self._lib.crane_sched_pending_erase(self._h, r.ctypes.data, len(r)))

    def set_running(self, running: abi.Running):
        c = running.as_c()
        self._check(self._lib.crane_sched_set_running(self._h, C.byref(c)))

    def pending_rows(self) -> int:
        return int(self._lib.crane_sched_pending_rows(self._h))

    def run(self, now: int):
        self._check(self._lib.crane_sched_run(self._h, now))

    def sync(self) -> float:
        """Waits for the handle's stream; returns the device ms of the last run."""
        ms = C.c_float(0.0)
        self._check(self._lib.crane_sched_sync(self._h, C.byref(ms)))
        return ms.value

    def fetch(self, out: abi.Placements) -> abi.Placements:
        c_out = out.as_c()
        self._check(self._lib.crane_sched_fetch(self._h, C.byref(c_out)))
        return out

    # --- one queue over several GPUs (include/crane_sched.h: crane_sched_set_shard) ---
    def set_shard(self, rank: int, n_ranks: int, part_owner):
        import numpy as np
        owner = np.ascontiguousarray(part_owner, np.uint32) if part_owner is not None else None
        self._check(self._lib.crane_sched_set_shard(self._h, rank, n_ranks, owner.ctypes.data if owner is not None else None))

    def device_placements(self) -> "abi.DevicePlacementsC":
        d = abi.DevicePlacementsC()
        self._check(self._lib.crane_sched_device_placements(self._h, C.byref(d)))
        return d

    def qos_filter(self, qos: abi.QosTable, reason: np.ndarray) -> np.ndarray:
        """CheckAndMallocQosResource pass of the commit loop (JobScheduler.cpp:1262)
        over the placements of the last run; updates `reason` and qos.*_usage."""
        c_q = qos.as_c()
        self._check(self._lib.crane_sched_qos_filter(self._h, C.byref(c_q), reason.ctypes.data))
        return reason

    def timing(self) -> dict:
        t = abi.TimingC()
        self._check(self._lib.crane_sched_get_timing(self._h, C.byref(t)))
        return {k: getattr(t, k) for k, _ in abi.TimingC._fields_}

    def debug_profile(self):
        import numpy as np
        buf = np.zeros((self.cluster.n_partitions, 16), np.uint64)
        self._check(self._lib.crane_sched_debug_profile(self._h, buf.ctypes.data, buf.size))
        return buf

    def debug_bitmap(self):
        import numpy as np
        rows, wpr = C.c_uint32(0), C.c_uint32(0)
        self._check(self._lib.crane_sched_debug_bitmap(self._h, None, 0, C.byref(rows), C.byref(wpr)))
        buf = np.zeros((rows.value, wpr.value), np.uint32)
        if buf.size:
            self._check(self._lib.crane_sched_debug_bitmap(self._h, buf.ctypes.data, buf.size, C.byref(rows), C.byref(wpr)))
        return buf
