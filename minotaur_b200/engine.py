"""Python host binding of the C ABI in include/mntr_gpu.h (libmntr_gpu.so), via ctypes.

This is plumbing for tests, bench.py and multi-GPU drivers: it owns no algorithm.  It fails
loudly when the CUDA library is missing or no CUDA device is usable -- there is no CPU path.

Method names follow the reference handler they stand in for:
``simple_presolve`` <-> LinearHandler::simplePresolve / NlPresHandler::simplePresolve
(/root/reference/src/base/LinearHandler.cpp:1605-1653, NlPresHandler.cpp:1022-1059),
``presolve_node`` <-> Handler::presolveNode (Handler.h:229-231; returns True = infeasible).
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass
from typing import Optional, Tuple

import numpy as np

from .instances import LinearRows, Tapes

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libmntr_gpu.so")

# enums of include/mntr_gpu.h
ROUND_DIRECTED, ROUND_NEAREST = 0, 1
ORDER_JACOBI, ORDER_REFERENCE, ORDER_AUTO = 0, 1, -1
LOOP_FIXPOINT, LOOP_SIMPLEPRESOLVE = 0, 1
FEASIBLE, INFEAS_BOUNDS, INFEAS_ROW, INFEAS_NL, ERROR_NL = 0, 1, 2, 3, 4
HANDLERS_ALL, HANDLERS_LINEAR, HANDLERS_NONLINEAR = 0, 1, 2
FLAG_PER_ROUND_KERNELS = 1
FLAG_STAGED_ROWS = 2

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)
_lp = C.POINTER(C.c_int64)
_bp = C.POINTER(C.c_uint8)

ABI_SYMBOLS = [
    "mntr_gpu_create", "mntr_gpu_destroy", "mntr_gpu_last_error", "mntr_gpu_abi_version",
    "mntr_gpu_device_count", "mntr_gpu_load_linear", "mntr_gpu_load_cgraph", "mntr_gpu_set_cutoff",
    "mntr_gpu_set_incumbent",
    "mntr_gpu_tighten", "mntr_gpu_tighten_nodes", "mntr_gpu_box_ld", "mntr_gpu_tighten_dev",
    "mntr_gpu_boxes_upload", "mntr_gpu_boxes_download", "mntr_gpu_get_stats",
    "mntr_gpu_tighten_single_dev", "mntr_gpu_stream",
    "mntr_gpu_nccl_unique_id", "mntr_gpu_comm_init", "mntr_gpu_comm_destroy",
    "mntr_gpu_boxes_from_deltas", "mntr_gpu_alloc_host", "mntr_gpu_free_host", "mntr_gpu_update_row_bounds",
    "mntr_gpu_load_quad", "mntr_gpu_group_load_quad", "mntr_gpu_root_dup_rows", "mntr_gpu_root_redundant_rows", "mntr_gpu_root_coeff_imp", "mntr_gpu_load_quad_relations", "mntr_gpu_quad_simple_presolve", "mntr_gpu_quad_presolve_node",
    "mntr_gpu_group_create", "mntr_gpu_group_destroy", "mntr_gpu_group_size", "mntr_gpu_group_member",
    "mntr_gpu_group_last_error", "mntr_gpu_group_load_linear", "mntr_gpu_group_load_cgraph",
    "mntr_gpu_group_set_cutoff", "mntr_gpu_group_set_incumbent", "mntr_gpu_group_tighten_nodes",
]


class GpuOptions(C.Structure):
    _fields_ = [("rounding", C.c_int32), ("order", C.c_int32), ("loop", C.c_int32), ("max_rounds", C.c_int32),
                ("handlers", C.c_int32), ("flags", C.c_int32), ("reserved", C.c_int32 * 2)]


class GpuStats(C.Structure):
    _fields_ = [("nnz_updates", C.c_int64), ("rows_evaluated", C.c_int64), ("n_infeasible", C.c_int64),
                ("n_changes", C.c_int64), ("max_rounds", C.c_int32), ("sparse_rounds", C.c_int32), ("kernel_ms", C.c_double),
                ("h2d_ms", C.c_double), ("d2h_ms", C.c_double), ("comm_ms", C.c_double), ("rows_ms", C.c_double),
                ("vars_ms", C.c_double), ("nl_evals", C.c_int64)]


class EngineError(RuntimeError):
    pass


_lib = None


def load_library() -> C.CDLL:
    """dlopen libmntr_gpu.so and declare the ABI.  Raises if the library is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise EngineError(f"{LIB_PATH} is missing: run `python -m minotaur_b200.build` "
                          "(there is no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    vp = C.c_void_p
    L.mntr_gpu_create.argtypes = [C.c_int, C.POINTER(vp)]
    L.mntr_gpu_destroy.argtypes = [vp]
    L.mntr_gpu_destroy.restype = None
    L.mntr_gpu_last_error.argtypes = [vp]
    L.mntr_gpu_last_error.restype = C.c_char_p
    L.mntr_gpu_load_linear.argtypes = [vp, C.c_int32, C.c_int32, _ip, _ip, _dp, _dp, _dp, _bp, _bp]
    L.mntr_gpu_load_cgraph.argtypes = [vp, C.c_int32, _ip, _bp, _ip, _ip, _dp, _ip, _ip, _ip, _dp, _dp, _dp]
    L.mntr_gpu_set_cutoff.argtypes = [vp, C.c_int32, _ip, _dp, C.c_double]
    L.mntr_gpu_set_incumbent.argtypes = [vp, C.c_double]
    L.mntr_gpu_tighten.argtypes = [vp, C.c_int32, _dp, _dp, C.POINTER(GpuOptions), _ip, _ip, _lp]
    L.mntr_gpu_tighten_nodes.argtypes = [vp, C.c_int32, _dp, _dp, _lp, _ip, _bp, _dp, C.POINTER(GpuOptions),
                                         _ip, _ip, _lp, _ip, _bp, _dp, C.c_int64, _lp]
    L.mntr_gpu_box_ld.argtypes = [C.c_int32]
    L.mntr_gpu_box_ld.restype = C.c_int64
    L.mntr_gpu_tighten_dev.argtypes = [vp, C.c_int32, vp, C.POINTER(GpuOptions), vp, vp, vp]
    L.mntr_gpu_boxes_upload.argtypes = [vp, C.c_int32, _dp, _dp, vp]
    L.mntr_gpu_boxes_download.argtypes = [vp, C.c_int32, vp, _dp, _dp]
    L.mntr_gpu_get_stats.argtypes = [vp, C.POINTER(GpuStats)]
    L.mntr_gpu_tighten_single_dev.argtypes = [vp, vp, vp, C.POINTER(GpuOptions), _ip, _ip, _lp]
    L.mntr_gpu_stream.argtypes = [vp]
    L.mntr_gpu_stream.restype = vp
    L.mntr_gpu_nccl_unique_id.argtypes = [vp]
    L.mntr_gpu_comm_init.argtypes = [vp, C.c_int32, C.c_int32, vp]
    L.mntr_gpu_comm_destroy.argtypes = [vp]
    L.mntr_gpu_boxes_from_deltas.argtypes = [vp, C.c_int32, _dp, _dp, _lp, _ip, _bp, _dp, vp]
    L.mntr_gpu_update_row_bounds.argtypes = [vp, C.c_int32, _dp, _dp]
    L.mntr_gpu_load_quad.argtypes = [vp, C.c_int32, _ip, _ip, _ip, _dp, _ip, _ip, _dp, _dp, _dp]
    L.mntr_gpu_root_dup_rows.argtypes = [vp, _dp, _dp, _dp, _dp, C.c_int64, _ip, _ip, _bp, _lp]
    L.mntr_gpu_root_redundant_rows.argtypes = [vp, _dp, _dp, _bp, _lp]
    L.mntr_gpu_load_quad_relations.argtypes = [vp, C.c_int32, _ip, _ip, C.c_int32, _ip, _ip, _ip]
    L.mntr_gpu_quad_simple_presolve.argtypes = [vp, C.c_int32, _dp, _dp, C.c_int32, _ip, _ip]
    L.mntr_gpu_quad_presolve_node.argtypes = [vp, C.c_int32, _dp, _dp, C.c_int32, C.c_int32, _ip, _ip, _ip]
    L.mntr_gpu_root_coeff_imp.argtypes = [vp, C.c_int32, C.c_int32, _ip, _ip, _dp, _dp, _dp, _bp, _dp, _dp, C.c_int64, _ip, _ip, _dp,
                                          _ip, _dp, _dp, _lp, _ip, _ip]
    L.mntr_gpu_group_load_quad.argtypes = [vp, C.c_int32, _ip, _ip, _ip, _dp, _ip, _ip, _dp, _dp, _dp]
    L.mntr_gpu_alloc_host.argtypes = [vp, C.c_int64]
    L.mntr_gpu_alloc_host.restype = vp
    L.mntr_gpu_free_host.argtypes = [vp, vp]
    L.mntr_gpu_free_host.restype = None
    L.mntr_gpu_group_create.argtypes = [C.c_int32, _ip, C.POINTER(vp)]
    L.mntr_gpu_group_destroy.argtypes = [vp]
    L.mntr_gpu_group_destroy.restype = None
    L.mntr_gpu_group_size.argtypes = [vp]
    L.mntr_gpu_group_member.argtypes = [vp, C.c_int32]
    L.mntr_gpu_group_member.restype = vp
    L.mntr_gpu_group_last_error.argtypes = [vp]
    L.mntr_gpu_group_last_error.restype = C.c_char_p
    L.mntr_gpu_group_load_linear.argtypes = [vp, C.c_int32, C.c_int32, _ip, _ip, _dp, _dp, _dp, _bp, _bp]
    L.mntr_gpu_group_load_cgraph.argtypes = [vp, C.c_int32, _ip, _bp, _ip, _ip, _dp, _ip, _ip, _ip, _dp, _dp, _dp]
    L.mntr_gpu_group_set_cutoff.argtypes = [vp, C.c_int32, _ip, _dp, C.c_double]
    L.mntr_gpu_group_set_incumbent.argtypes = [vp, C.c_double]
    L.mntr_gpu_group_tighten_nodes.argtypes = [vp, C.c_int32, _dp, _dp, _lp, _ip, _bp, _dp, C.POINTER(GpuOptions),
                                               _ip, _ip, _lp, _ip, _bp, _dp, C.c_int64, _lp]
    _lib = L
    return L


def _d(a): return a.ctypes.data_as(_dp)
def _i(a): return a.ctypes.data_as(_ip)
def _l(a): return a.ctypes.data_as(_lp)
def _b(a): return a.ctypes.data_as(_bp)


@dataclass
class TightenResult:
    lb: np.ndarray
    ub: np.ndarray
    verdict: np.ndarray       # int32 per box
    rounds: np.ndarray        # int32 per box
    nnz_updates: np.ndarray   # int64 per box
    kernel_ms: float = 0.0
    h2d_ms: float = 0.0
    d2h_ms: float = 0.0


def _tighten_nodes(fn, handle, check, what, root_lb, root_ub, delta_ptr, delta_var, delta_is_upper, delta_val, rounding,
                   loop, max_rounds, handlers, mod_cap, out=None):
    rl = np.ascontiguousarray(root_lb, np.float64); ru = np.ascontiguousarray(root_ub, np.float64)
    dp = np.ascontiguousarray(delta_ptr, np.int64); dv = np.ascontiguousarray(delta_var, np.int32)
    du = np.ascontiguousarray(delta_is_upper, np.uint8); dx = np.ascontiguousarray(delta_val, np.float64)
    nb = len(dp) - 1
    if len(dv) == 0:
        dv = np.zeros(1, np.int32); du = np.zeros(1, np.uint8); dx = np.zeros(1)
    o = GpuOptions(rounding, ORDER_REFERENCE, loop, max_rounds, handlers)
    v = np.zeros(nb, np.int32); r = np.zeros(nb, np.int32); mp = np.zeros(nb + 1, np.int64)
    cap = int(mod_cap) if mod_cap is not None else max(1024, 4 * nb)
    while True:
        if out is not None:       # caller-owned (e.g. page-locked) output buffers of at least mod_cap entries
            mv, mu, mx = out
            assert mod_cap is not None and len(mv) >= cap and len(mu) >= cap and len(mx) >= cap
        else:
            mv = np.zeros(max(cap, 1), np.int32); mu = np.zeros(max(cap, 1), np.uint8); mx = np.zeros(max(cap, 1))
        total = C.c_int64(0)
        check(fn(handle, nb, _d(rl), _d(ru), _l(dp), _i(dv), _b(du), _d(dx), C.byref(o),
                 _i(v), _i(r), _l(mp), _i(mv), _b(mu), _d(mx), cap, C.byref(total)), what)
        if total.value <= cap or mod_cap is not None:
            break
        cap = int(total.value)             # the buffer was too small: call again (documented contract)
    k = min(int(total.value), cap)
    return v, r, mp, mv[:k], mu[:k], mx[:k], int(total.value)


class GpuBoundEngine:
    """One engine context = one handler instance of the reference (one per B&B thread)."""

    def __init__(self, device: int = 0):
        self.L = load_library()
        self.h = C.c_void_p()
        rc = self.L.mntr_gpu_create(device, C.byref(self.h))
        if rc != 0:
            raise EngineError(f"mntr_gpu_create(device={device}) failed with {rc}: no usable CUDA device "
                              "(the engine has no CPU fallback)")
        self.n = 0
        self.m = 0
        self.nnz = 0

    # -- lifetime --
    def close(self):
        if getattr(self, "h", None) is not None and self.h:
            self.L.mntr_gpu_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _check(self, rc: int, what: str):
        if rc != 0:
            msg = self.L.mntr_gpu_last_error(self.h)
            raise EngineError(f"{what} failed ({rc}): {msg.decode() if msg else ''}")

    # -- upload --
    def load_linear(self, inst: LinearRows):
        rp = np.ascontiguousarray(inst.row_ptr, np.int32); col = np.ascontiguousarray(inst.col, np.int32)
        val = np.ascontiguousarray(inst.val, np.float64)
        rl = np.ascontiguousarray(inst.row_lb, np.float64); ru = np.ascontiguousarray(inst.row_ub, np.float64)
        vt = np.ascontiguousarray(inst.var_type, np.uint8)
        ra = None if inst.row_active is None else np.ascontiguousarray(inst.row_active, np.uint8)
        self._check(self.L.mntr_gpu_load_linear(self.h, inst.m, inst.n, _i(rp), _i(col), _d(val), _d(rl), _d(ru),
                                                _b(vt), _b(ra) if ra is not None else None), "load_linear")
        self.n, self.m, self.nnz = inst.n, inst.m, inst.nnz
        if inst.cut_col is not None and len(inst.cut_col):
            self.set_cutoff(inst.cut_col, inst.cut_val, inst.cut_rhs)
            self.set_incumbent(inst.cut_rhs + getattr(inst, "obj_const", 0.0))

    def load_cgraph(self, t: Tapes):
        a = {k: np.ascontiguousarray(getattr(t, k), ty) for k, ty in (
            ("tape_ptr", np.int32), ("op", np.uint8), ("arg0", np.int32), ("arg1", np.int32), ("cnst", np.float64),
            ("child", np.int32), ("lin_ptr", np.int32), ("lin_col", np.int32), ("lin_val", np.float64),
            ("c_lb", np.float64), ("c_ub", np.float64))}
        self._check(self.L.mntr_gpu_load_cgraph(self.h, t.n_cons, _i(a["tape_ptr"]), _b(a["op"]), _i(a["arg0"]),
                                                _i(a["arg1"]), _d(a["cnst"]), _i(a["child"]), _i(a["lin_ptr"]),
                                                _i(a["lin_col"]), _d(a["lin_val"]), _d(a["c_lb"]), _d(a["c_ub"])),
                    "load_cgraph")

    def load_quad(self, q):
        """QuadraticFunction constraints (instances.QuadCons); None or an empty set removes them."""
        if q is None or q.n_quad == 0:
            self._check(self.L.mntr_gpu_load_quad(self.h, 0, None, None, None, None, None, None, None, None, None), "load_quad")
            return
        a = {k: np.ascontiguousarray(getattr(q, k), ty) for k, ty in (
            ("q_ptr", np.int32), ("v1", np.int32), ("v2", np.int32), ("coef", np.float64), ("lin_ptr", np.int32),
            ("lin_col", np.int32), ("lin_val", np.float64), ("q_lb", np.float64), ("q_ub", np.float64))}
        self._check(self.L.mntr_gpu_load_quad(self.h, q.n_quad, _i(a["q_ptr"]), _i(a["v1"]), _i(a["v2"]), _d(a["coef"]),
                                              _i(a["lin_ptr"]), _i(a["lin_col"]), _d(a["lin_val"]), _d(a["q_lb"]), _d(a["q_ub"])),
                    "load_quad")

    def root_dup_rows(self, r1, r2, cap=1 << 16):
        """Duplicate-row candidates of LinearHandler::dupRows_: returns (h1, h2, pairs) with pairs an int array [k, 3]
        of (i, j, kind) sorted by (i, j)."""
        r1 = np.ascontiguousarray(r1, np.float64); r2 = np.ascontiguousarray(r2, np.float64)
        h1 = np.zeros(max(self.m, 1)); h2 = np.zeros(max(self.m, 1))
        while True:
            pi = np.zeros(max(cap, 1), np.int32); pj = np.zeros(max(cap, 1), np.int32); pk = np.zeros(max(cap, 1), np.uint8)
            total = C.c_int64(0)
            self._check(self.L.mntr_gpu_root_dup_rows(self.h, _d(r1), _d(r2), _d(h1), _d(h2), cap, _i(pi), _i(pj), _b(pk),
                                                      C.byref(total)), "root_dup_rows")
            if total.value <= cap:
                break
            cap = int(total.value)
        k = int(total.value)
        return h1[:self.m], h2[:self.m], np.stack([pi[:k], pj[:k], pk[:k].astype(np.int32)], axis=1)

    def root_redundant_rows(self, lb, ub):
        """Rows whose activity range lies inside their bounds on the box (linBndTighten_ root mode, :974-985)."""
        lb = np.ascontiguousarray(lb, np.float64); ub = np.ascontiguousarray(ub, np.float64)
        out = np.zeros(max(self.m, 1), np.uint8); cnt = C.c_int64(0)
        self._check(self.L.mntr_gpu_root_redundant_rows(self.h, _d(lb), _d(ub), _b(out), C.byref(cnt)), "root_redundant_rows")
        return out[:self.m].astype(bool)

    def root_coeff_imp(self, inst, lb, ub, cap=1 << 20):
        """LinearHandler::coeffImp_ (LinearHandler.cpp:600-704) on the device for the rows of ``inst`` (caller's order; nothing
        needs to be loaded): (row, var, new coefficient, side, new row bound) arrays sorted by row, plus
        {'levels', 'erased', 'kernel_ms'}."""
        rp = np.ascontiguousarray(inst.row_ptr, np.int32); col = np.ascontiguousarray(inst.col, np.int32)
        val = np.ascontiguousarray(inst.val, np.float64)
        rl = np.ascontiguousarray(inst.row_lb, np.float64); ru = np.ascontiguousarray(inst.row_ub, np.float64)
        vt = np.ascontiguousarray(inst.var_type, np.uint8)
        lb = np.ascontiguousarray(lb, np.float64); ub = np.ascontiguousarray(ub, np.float64)
        row = np.zeros(max(cap, 1), np.int32); var = np.zeros(max(cap, 1), np.int32); coef = np.zeros(max(cap, 1))
        side = np.zeros(max(cap, 1), np.int32); bnd = np.zeros(max(cap, 1))
        cnt = C.c_int64(0); lev = C.c_int32(0); er = C.c_int32(0)
        self._check(self.L.mntr_gpu_root_coeff_imp(self.h, inst.m, inst.n, _i(rp), _i(col), _d(val), _d(rl), _d(ru), _b(vt), _d(lb),
                                                   _d(ub), cap, _i(row), _i(var), _d(coef), _i(side), _d(bnd), None, C.byref(cnt),
                                                   C.byref(lev), C.byref(er)), "root_coeff_imp")
        k = min(int(cnt.value), cap)
        info = {"levels": int(lev.value), "erased": int(er.value), "count": int(cnt.value), "kernel_ms": self.stats().kernel_ms}
        return row[:k], var[:k], coef[:k], side[:k], bnd[:k], info

    def load_quad_relations(self, rel):
        """The relations of a QuadHandler (instances.QuadRelations, in the handler's container order)."""
        a = [np.ascontiguousarray(x if len(x) else np.zeros(1), np.int32) for x in (rel.sq_x, rel.sq_y, rel.b_x0, rel.b_x1, rel.b_y)]
        self._check(self.L.mntr_gpu_load_quad_relations(self.h, len(rel.sq_x), _i(a[0]), _i(a[1]), len(rel.b_x0), _i(a[2]), _i(a[3]),
                                                        _i(a[4])), "load_quad_relations")

    def quad_simple_presolve(self, lb, ub, rounding=ROUND_NEAREST):
        """QuadHandler::simplePresolve on boxes [n_boxes][n]: (lb, ub, n_mods, n_inconsistent, kernel_ms)."""
        lb = np.array(np.atleast_2d(lb), np.float64, order="C"); ub = np.array(np.atleast_2d(ub), np.float64, order="C")
        nb = lb.shape[0]
        nm = np.zeros(nb, np.int32); bad = np.zeros(nb, np.int32)
        self._check(self.L.mntr_gpu_quad_simple_presolve(self.h, nb, _d(lb), _d(ub), rounding, _i(nm), _i(bad)), "quad_simple_presolve")
        return lb, ub, nm, bad, self.stats().kernel_ms

    def quad_presolve_node(self, lb, ub, rounding=ROUND_NEAREST, max_sweeps=0):
        """The propagation loop of QuadHandler::presolveNode on boxes [n_boxes][n]:
        (lb, ub, verdict, n_mods, n_sweeps, kernel_ms)."""
        lb = np.array(np.atleast_2d(lb), np.float64, order="C"); ub = np.array(np.atleast_2d(ub), np.float64, order="C")
        nb = lb.shape[0]
        v = np.zeros(nb, np.int32); nm = np.zeros(nb, np.int32); ns = np.zeros(nb, np.int32)
        self._check(self.L.mntr_gpu_quad_presolve_node(self.h, nb, _d(lb), _d(ub), rounding, max_sweeps, _i(v), _i(nm), _i(ns)),
                    "quad_presolve_node")
        return lb, ub, v, nm, ns, self.stats().kernel_ms

    def update_row_bounds(self, row_lb, row_ub):
        """New bounds for the loaded rows (same order as load_linear): m doubles each way, no re-flattening."""
        rl = np.ascontiguousarray(row_lb, np.float64); ru = np.ascontiguousarray(row_ub, np.float64)
        self._check(self.L.mntr_gpu_update_row_bounds(self.h, len(rl), _d(rl), _d(ru)), "update_row_bounds")

    def set_cutoff(self, col, val, rhs: float):
        col = np.ascontiguousarray(col, np.int32); val = np.ascontiguousarray(val, np.float64)
        self._check(self.L.mntr_gpu_set_cutoff(self.h, len(col), _i(col), _d(val), float(rhs)), "set_cutoff")

    def set_incumbent(self, best_value: float):
        """Raw incumbent value for NlPresHandler::fixObjBins_ (needs a cut-off row; inf switches the rule off)."""
        self._check(self.L.mntr_gpu_set_incumbent(self.h, float(best_value)), "set_incumbent")

    # -- the hot path --
    def tighten(self, lb, ub, rounding=ROUND_DIRECTED, order=ORDER_AUTO, loop=LOOP_FIXPOINT, max_rounds=0,
                inplace=False, handlers=HANDLERS_ALL, flags=0) -> TightenResult:
        """lb/ub: [n] or box-major [n_boxes, n] float64 host arrays."""
        lb = np.asarray(lb, np.float64); ub = np.asarray(ub, np.float64)
        if not inplace or not lb.flags.c_contiguous or not ub.flags.c_contiguous:
            lb = np.array(lb, np.float64, order="C", copy=True); ub = np.array(ub, np.float64, order="C", copy=True)
        single = lb.ndim == 1
        nb = 1 if single else lb.shape[0]
        if lb.shape[-1] != self.n or ub.shape != lb.shape:
            raise ValueError("box shape does not match the loaded problem")
        o = GpuOptions(rounding, order, loop, max_rounds, handlers, flags)
        v = np.zeros(nb, np.int32); r = np.zeros(nb, np.int32); z = np.zeros(nb, np.int64)
        self._check(self.L.mntr_gpu_tighten(self.h, nb, _d(lb), _d(ub), C.byref(o), _i(v), _i(r), _l(z)), "tighten")
        st = self.stats()
        return TightenResult(lb, ub, v, r, z, st.kernel_ms, st.h2d_ms, st.d2h_ms)

    def tighten_raw(self, nb: int, lb_ptr: int, ub_ptr: int, opts: GpuOptions, v_ptr: int = 0, r_ptr: int = 0,
                    z_ptr: int = 0):
        """Pointer-level call for callers that own (pinned) host buffers."""
        self._check(self.L.mntr_gpu_tighten(self.h, nb, C.cast(lb_ptr, _dp), C.cast(ub_ptr, _dp), C.byref(opts),
                                            C.cast(v_ptr, _ip), C.cast(r_ptr, _ip), C.cast(z_ptr, _lp)), "tighten")

    def simple_presolve(self, lb, ub, **kw) -> TightenResult:
        """LinearHandler::simplePresolve semantics (loop truncation of the reference)."""
        kw.setdefault("loop", LOOP_SIMPLEPRESOLVE)
        return self.tighten(lb, ub, **kw)

    def presolve_node(self, lb, ub, **kw) -> Tuple[bool, TightenResult]:
        """Handler::presolveNode: returns (is_infeasible, result)."""
        res = self.simple_presolve(lb, ub, **kw)
        return bool(res.verdict[0] != FEASIBLE), res

    # -- device-resident boxes --
    def box_ld(self, n_boxes: int) -> int:
        return int(self.L.mntr_gpu_box_ld(n_boxes))

    def boxes_upload(self, lb, ub, boxes_dev_ptr: int):
        lb = np.ascontiguousarray(lb, np.float64); ub = np.ascontiguousarray(ub, np.float64)
        self._check(self.L.mntr_gpu_boxes_upload(self.h, lb.shape[0], _d(lb), _d(ub), C.c_void_p(boxes_dev_ptr)),
                    "boxes_upload")

    def boxes_download(self, n_boxes: int, boxes_dev_ptr: int):
        lb = np.zeros((n_boxes, self.n)); ub = np.zeros((n_boxes, self.n))
        self._check(self.L.mntr_gpu_boxes_download(self.h, n_boxes, C.c_void_p(boxes_dev_ptr), _d(lb), _d(ub)),
                    "boxes_download")
        return lb, ub

    def tighten_nodes(self, root_lb, root_ub, delta_ptr, delta_var, delta_is_upper, delta_val, rounding=ROUND_DIRECTED,
                      loop=LOOP_FIXPOINT, max_rounds=0, handlers=HANDLERS_ALL, mod_cap=None, out=None):
        """Node batch as branching deltas on a root box; returns (verdict, rounds, mod_ptr, mod_var, mod_is_upper,
        mod_val): the VarBoundMod tuples of every feasible box, ascending (variable, side) inside a box."""
        return _tighten_nodes(self.L.mntr_gpu_tighten_nodes, self.h, self._check, "tighten_nodes", root_lb, root_ub,
                              delta_ptr, delta_var, delta_is_upper, delta_val, rounding, loop, max_rounds, handlers, mod_cap,
                              out)

    def alloc_host_bytes(self, nbytes: int) -> np.ndarray:
        """`nbytes` of page-locked, device-mapped host memory as a uint8 array (view it as any dtype)."""
        p = self.L.mntr_gpu_alloc_host(self.h, int(nbytes))
        if not p:
            raise EngineError("mntr_gpu_alloc_host failed")
        return np.ctypeslib.as_array(C.cast(p, _bp), shape=(int(nbytes),))

    def boxes_from_deltas(self, root_lb, root_ub, delta_ptr, delta_var, delta_is_upper, delta_val, boxes_dev_ptr: int):
        """Builds the node boxes (root + deltas) in the engine's layout on the device."""
        rl = np.ascontiguousarray(root_lb, np.float64); ru = np.ascontiguousarray(root_ub, np.float64)
        dp = np.ascontiguousarray(delta_ptr, np.int64); dv = np.ascontiguousarray(delta_var, np.int32)
        du = np.ascontiguousarray(delta_is_upper, np.uint8); dx = np.ascontiguousarray(delta_val, np.float64)
        if len(dv) == 0:
            dv = np.zeros(1, np.int32); du = np.zeros(1, np.uint8); dx = np.zeros(1)
        self._check(self.L.mntr_gpu_boxes_from_deltas(self.h, len(dp) - 1, _d(rl), _d(ru), _l(dp), _i(dv), _b(du), _d(dx),
                                                      C.c_void_p(boxes_dev_ptr)), "boxes_from_deltas")

    def alloc_host(self, count: int) -> np.ndarray:
        """`count` float64 in page-locked, device-mapped host memory (mntr_gpu_alloc_host): a single-box tighten on
        such arrays takes the zero-copy path.  Free with free_host()."""
        p = self.L.mntr_gpu_alloc_host(self.h, 8 * int(count))
        if not p:
            raise EngineError("mntr_gpu_alloc_host failed")
        a = np.ctypeslib.as_array(C.cast(p, _dp), shape=(int(count),))
        return a

    def free_host(self, a: np.ndarray):
        self.L.mntr_gpu_free_host(self.h, C.c_void_p(a.ctypes.data))

    def tighten_dev(self, n_boxes: int, boxes_dev_ptr: int, verdict_ptr: int, rounds_ptr: int, nnz_ptr: int,
                    rounding=ROUND_DIRECTED, loop=LOOP_FIXPOINT, max_rounds=0, handlers=HANDLERS_ALL):
        o = GpuOptions(rounding, ORDER_REFERENCE, loop, max_rounds, handlers)
        self._check(self.L.mntr_gpu_tighten_dev(self.h, n_boxes, C.c_void_p(boxes_dev_ptr), C.byref(o),
                                                C.c_void_p(verdict_ptr), C.c_void_p(rounds_ptr),
                                                C.c_void_p(nnz_ptr)), "tighten_dev")
        return self.stats()

    def tighten_single_dev(self, lb_dev_ptr: int, ub_dev_ptr: int, rounding=ROUND_DIRECTED, loop=LOOP_FIXPOINT,
                           max_rounds=0, flags=0):
        """Single box resident in HBM (device pointers), Jacobi fixpoint; returns (verdict, rounds, nnz)."""
        o = GpuOptions(rounding, ORDER_JACOBI, loop, max_rounds, HANDLERS_ALL, flags)
        v = C.c_int32(0); r = C.c_int32(0); z = C.c_int64(0)
        self._check(self.L.mntr_gpu_tighten_single_dev(self.h, C.c_void_p(lb_dev_ptr), C.c_void_p(ub_dev_ptr),
                                                       C.byref(o), C.byref(v), C.byref(r), C.byref(z)),
                    "tighten_single_dev")
        return v.value, r.value, z.value

    # -- row-partitioned multi-GPU mode --
    @staticmethod
    def nccl_unique_id() -> bytes:
        """128-byte NCCL unique id (create on rank 0, hand to every rank's comm_init)."""
        L = load_library()
        buf = C.create_string_buffer(128)
        rc = L.mntr_gpu_nccl_unique_id(buf)
        if rc != 0:
            raise EngineError(f"mntr_gpu_nccl_unique_id failed ({rc}): NCCL unavailable")
        return buf.raw

    def comm_init(self, n_ranks: int, rank: int, unique_id: bytes):
        buf = C.create_string_buffer(unique_id, 128)
        self._check(self.L.mntr_gpu_comm_init(self.h, n_ranks, rank, buf), "comm_init")

    def comm_destroy(self):
        self._check(self.L.mntr_gpu_comm_destroy(self.h), "comm_destroy")

    def stream_handle(self) -> int:
        return int(self.L.mntr_gpu_stream(self.h) or 0)

    def stats(self) -> GpuStats:
        s = GpuStats()
        self._check(self.L.mntr_gpu_get_stats(self.h, C.byref(s)), "get_stats")
        return s


class GpuBoundGroup:
    """Several GPUs of one box driven from ONE process (mntr_gpu_group_*): the problem is replicated, a node batch is
    split contiguously over the devices."""

    def __init__(self, devices):
        self.L = load_library()
        self.h = C.c_void_p()
        dv = np.ascontiguousarray(devices, np.int32)
        rc = self.L.mntr_gpu_group_create(len(dv), _i(dv), C.byref(self.h))
        if rc != 0:
            raise EngineError(f"mntr_gpu_group_create({list(devices)}) failed with {rc} (there is no CPU fallback)")
        self.n = 0

    def close(self):
        if getattr(self, "h", None) is not None and self.h:
            self.L.mntr_gpu_group_destroy(self.h)
            self.h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int, what: str):
        if rc != 0:
            msg = self.L.mntr_gpu_group_last_error(self.h)
            raise EngineError(f"{what} failed ({rc}): {msg.decode() if msg else ''}")

    @property
    def size(self) -> int:
        return int(self.L.mntr_gpu_group_size(self.h))

    def load_linear(self, inst: LinearRows):
        rp = np.ascontiguousarray(inst.row_ptr, np.int32); col = np.ascontiguousarray(inst.col, np.int32)
        val = np.ascontiguousarray(inst.val, np.float64)
        rl = np.ascontiguousarray(inst.row_lb, np.float64); ru = np.ascontiguousarray(inst.row_ub, np.float64)
        vt = np.ascontiguousarray(inst.var_type, np.uint8)
        ra = None if inst.row_active is None else np.ascontiguousarray(inst.row_active, np.uint8)
        self._check(self.L.mntr_gpu_group_load_linear(self.h, inst.m, inst.n, _i(rp), _i(col), _d(val), _d(rl), _d(ru),
                                                      _b(vt), _b(ra) if ra is not None else None), "group_load_linear")
        self.n = inst.n
        if inst.cut_col is not None and len(inst.cut_col):
            cc = np.ascontiguousarray(inst.cut_col, np.int32); cv = np.ascontiguousarray(inst.cut_val, np.float64)
            self._check(self.L.mntr_gpu_group_set_cutoff(self.h, len(cc), _i(cc), _d(cv), float(inst.cut_rhs)), "group_set_cutoff")
            self._check(self.L.mntr_gpu_group_set_incumbent(self.h, float(inst.cut_rhs + getattr(inst, "obj_const", 0.0))),
                        "group_set_incumbent")

    def load_cgraph(self, t: Tapes):
        a = {k: np.ascontiguousarray(getattr(t, k), ty) for k, ty in (
            ("tape_ptr", np.int32), ("op", np.uint8), ("arg0", np.int32), ("arg1", np.int32), ("cnst", np.float64),
            ("child", np.int32), ("lin_ptr", np.int32), ("lin_col", np.int32), ("lin_val", np.float64),
            ("c_lb", np.float64), ("c_ub", np.float64))}
        self._check(self.L.mntr_gpu_group_load_cgraph(self.h, t.n_cons, _i(a["tape_ptr"]), _b(a["op"]), _i(a["arg0"]),
                                                      _i(a["arg1"]), _d(a["cnst"]), _i(a["child"]), _i(a["lin_ptr"]),
                                                      _i(a["lin_col"]), _d(a["lin_val"]), _d(a["c_lb"]), _d(a["c_ub"])),
                    "group_load_cgraph")

    def tighten_nodes(self, root_lb, root_ub, delta_ptr, delta_var, delta_is_upper, delta_val, rounding=ROUND_DIRECTED,
                      loop=LOOP_FIXPOINT, max_rounds=0, handlers=HANDLERS_ALL, mod_cap=None):
        return _tighten_nodes(self.L.mntr_gpu_group_tighten_nodes, self.h, self._check, "group_tighten_nodes", root_lb,
                              root_ub, delta_ptr, delta_var, delta_is_upper, delta_val, rounding, loop, max_rounds, handlers,
                              mod_cap)
