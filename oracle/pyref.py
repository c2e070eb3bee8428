"""ctypes front-end of oracle/_ref/libcrane_ref.so — the reference's own
NodeSelect compiled from its source text (oracle/ref_build.py).
TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import numpy as np

from cranesched_b200 import abi
from oracle import ref_build

_lib = None
_p = C.c_void_p


class RefExtraC(C.Structure):
    _fields_ = [
        ("n_resv", C.c_uint32), ("resv_start", _p), ("resv_end", _p), ("resv_off", _p), ("resv_node", _p),
        ("resv_res", _p), ("pd_resv", _p), ("rn_resv", _p),
        ("n_qos", C.c_uint32), ("preempt_qos", _p), ("rn_qos", _p), ("preempted_running", _p),
    ]


@dataclass
class RefExtra:
    """Reservations / preemption inputs (crane_ref_extra_t)."""
    resv_start: np.ndarray | None = None
    resv_end: np.ndarray | None = None
    resv_off: np.ndarray | None = None
    resv_node: np.ndarray | None = None
    resv_res: np.ndarray | None = None
    pd_resv: np.ndarray | None = None
    rn_resv: np.ndarray | None = None
    preempt_qos: np.ndarray | None = None
    rn_qos: np.ndarray | None = None
    preempted_running: np.ndarray | None = None
    _keep: list = field(default_factory=list)

    def as_c(self) -> RefExtraC:
        def a(x, dt):
            if x is None:
                return None
            y = np.ascontiguousarray(x, dt)
            self._keep.append(y)
            return y.ctypes.data
        n_resv = 0 if self.resv_start is None else len(self.resv_start)
        n_qos = 0 if self.preempt_qos is None else int(np.asarray(self.preempt_qos).shape[0])
        return RefExtraC(n_resv, a(self.resv_start, np.int64), a(self.resv_end, np.int64), a(self.resv_off, np.uint32),
                         a(self.resv_node, np.uint32), a(self.resv_res, abi.RES_IN_NODE), a(self.pd_resv, np.uint32),
                         a(self.rn_resv, np.uint32), n_qos, a(self.preempt_qos, np.uint8), a(self.rn_qos, np.uint32),
                         None if self.preempted_running is None else self.preempted_running.ctypes.data)


def available() -> bool:
    return ref_build.build() is not None


def lib():
    global _lib
    if _lib is None:
        so = ref_build.build()
        if so is None:
            raise RuntimeError("oracle/_ref/libcrane_ref.so is not built and /root/reference is absent")
        _lib = C.CDLL(so)
        _lib.crane_ref_node_select.restype = C.c_int
        _lib.crane_ref_node_select.argtypes = [
            C.POINTER(abi.SchedConfig), C.POINTER(abi.ClusterC), C.c_int64, C.POINTER(abi.RunningC),
            C.POINTER(abi.PendingC), C.POINTER(RefExtraC), C.POINTER(abi.PlacementsC), C.POINTER(C.c_double)]
        _lib.crane_ref_feasible.restype = C.c_int
        _lib.crane_ref_feasible.argtypes = [C.POINTER(abi.ClusterC), _p, _p, _p]
        _lib.crane_ref_ckmin.restype = None
        _lib.crane_ref_ckmin.argtypes = [C.POINTER(abi.ClusterC), _p, _p]
        _lib.crane_ref_qos_filter.restype = C.c_int
        _lib.crane_ref_qos_filter.argtypes = [C.POINTER(abi.ClusterC), C.POINTER(abi.PendingC), C.POINTER(abi.PlacementsC),
                                              C.POINTER(abi.QosTableC)]
        _lib.crane_ref_res_le.restype = C.c_int
        _lib.crane_ref_res_le.argtypes = [C.POINTER(abi.ClusterC), _p, _p]
    return _lib


def node_select(cfg, cluster, running, pending, now, extra: RefExtra | None = None):
    """Returns (Placements, elapsed_ms) of the reference's own NodeSelect."""
    out = abi.Placements.for_pending(pending)
    c_cfg, c_cl, c_rn, c_pd, c_out = cfg.as_c(), cluster.as_c(), running.as_c(), pending.as_c(), out.as_c()
    ms = C.c_double(0.0)
    c_ex = extra.as_c() if extra is not None else None
    rc = lib().crane_ref_node_select(C.byref(c_cfg), C.byref(c_cl), now, C.byref(c_rn), C.byref(c_pd),
                                     C.byref(c_ex) if c_ex is not None else None, C.byref(c_out), C.byref(ms))
    if rc != 0:
        raise RuntimeError(f"crane_ref_node_select rc={rc}")
    return out, ms.value


def feasible(cluster, req, avail):
    alloc = np.zeros((), abi.RES_IN_NODE)
    req = np.ascontiguousarray(req, abi.RES_VIEW)
    avail = np.ascontiguousarray(avail, abi.RES_IN_NODE)
    c = cluster.as_c()
    ok = lib().crane_ref_feasible(C.byref(c), req.ctypes.data, avail.ctypes.data, alloc.ctypes.data)
    return bool(ok), alloc


def ckmin(cluster, a, b):
    a = np.array(a, abi.RES_IN_NODE, copy=True)
    b = np.ascontiguousarray(b, abi.RES_IN_NODE)
    c = cluster.as_c()
    lib().crane_ref_ckmin(C.byref(c), a.ctypes.data, b.ctypes.data)
    return a


def res_le(cluster, a, b) -> bool:
    a = np.ascontiguousarray(a, abi.RES_IN_NODE)
    b = np.ascontiguousarray(b, abi.RES_IN_NODE)
    c = cluster.as_c()
    return bool(lib().crane_ref_res_le(C.byref(c), a.ctypes.data, b.ctypes.data))


def qos_filter(cluster, pending, out, qos):
    """The reference's own CheckAndMallocQosResource (Accounting/AccountMetaContainer.cpp) over the jobs
    `out` starts now; in place on out.reason and qos.*_usage (same contract as pyoracle.qos_filter)."""
    c_cl, c_pd, c_out, c_q = cluster.as_c(), pending.as_c(), out.as_c(), qos.as_c()
    rc = lib().crane_ref_qos_filter(C.byref(c_cl), C.byref(c_pd), C.byref(c_out), C.byref(c_q))
    if rc != 0:
        raise RuntimeError(f"crane_ref_qos_filter rc={rc}")
    return out
