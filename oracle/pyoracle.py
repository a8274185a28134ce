"""ctypes bindings of the two CPU checkers.  TEST INFRASTRUCTURE ONLY.

* ``Oracle``    -> oracle/liboracle.so            (fbbt_oracle.c, our plain-C restatement)
* ``Reference`` -> oracle/_ref/libminotaur_ref.so (the reference's own objects + ref_harness.cpp)

Only tests/, ``__graft_entry__.smoke()`` and bench.py's ``cpu_baseline`` / ``--impl reference``
legs may import this module; nothing under ``minotaur_b200/`` does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Optional, Tuple

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "liboracle.so")
REF_SO = os.path.join(HERE, "_ref", "libminotaur_ref.so")

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)
_lp = C.POINTER(C.c_int64)
_bp = C.POINTER(C.c_uint8)


def _d(a): return a.ctypes.data_as(_dp)
def _i(a): return a.ctypes.data_as(_ip)
def _b(a): return a.ctypes.data_as(_bp)
def _l(a): return a.ctypes.data_as(_lp)


def _delta_arrays(deltas):
    ptr, var, up, val = deltas
    ptr = np.ascontiguousarray(ptr, np.int64); var = np.ascontiguousarray(var, np.int32)
    up = np.ascontiguousarray(up, np.uint8); val = np.ascontiguousarray(val, np.float64)
    if var.size == 0:
        var = np.zeros(1, np.int32); up = np.zeros(1, np.uint8); val = np.zeros(1)
    return ptr, var, up, val


class OrcLin(C.Structure):
    _fields_ = [("m", C.c_int32), ("n", C.c_int32), ("row_ptr", _ip), ("col", _ip), ("val", _dp),
                ("row_lb", _dp), ("row_ub", _dp), ("var_type", _bp), ("row_active", _bp),
                ("cut_k", C.c_int32), ("cut_col", _ip), ("cut_val", _dp), ("cut_rhs", C.c_double),
                ("obj_ub", C.c_double)]


class OrcNl(C.Structure):
    _fields_ = [("n_cons", C.c_int32), ("tape_ptr", _ip), ("op", _bp), ("arg0", _ip), ("arg1", _ip),
                ("cnst", _dp), ("child", _ip), ("lin_ptr", _ip), ("lin_col", _ip), ("lin_val", _dp),
                ("c_lb", _dp), ("c_ub", _dp),
                ("n_quad", C.c_int32), ("q_ptr", _ip), ("q_v1", _ip), ("q_v2", _ip), ("q_coef", _dp), ("q_lin_ptr", _ip),
                ("q_lin_col", _ip), ("q_lin_val", _dp), ("q_lb", _dp), ("q_ub", _dp)]


class OrcResult(C.Structure):
    _fields_ = [("verdict", C.c_int32), ("rounds", C.c_int32), ("nnz_updates", C.c_int64),
                ("n_mods", C.c_int64)]


def build_oracle(force: bool = False) -> str:
    """Compile oracle/liboracle.so (and oracle/_ref when /root/reference is present)."""
    if force or not os.path.exists(ORACLE_SO) or \
            os.path.getmtime(ORACLE_SO) < os.path.getmtime(os.path.join(HERE, "fbbt_oracle.c")):
        subprocess.check_call(["make", "-s", "-C", HERE, "oracle"])
    return ORACLE_SO


def have_reference() -> bool:
    return os.path.exists(REF_SO)


def _lin_struct(inst, keep):
    s = OrcLin()
    s.m, s.n = inst.m, inst.n
    arrs = dict(row_ptr=np.ascontiguousarray(inst.row_ptr, np.int32), col=np.ascontiguousarray(inst.col, np.int32),
                val=np.ascontiguousarray(inst.val, np.float64), row_lb=np.ascontiguousarray(inst.row_lb, np.float64),
                row_ub=np.ascontiguousarray(inst.row_ub, np.float64),
                var_type=np.ascontiguousarray(inst.var_type, np.uint8))
    keep.update(arrs)
    s.row_ptr, s.col, s.val = _i(arrs["row_ptr"]), _i(arrs["col"]), _d(arrs["val"])
    s.row_lb, s.row_ub, s.var_type = _d(arrs["row_lb"]), _d(arrs["row_ub"]), _b(arrs["var_type"])
    if inst.row_active is not None:
        keep["row_active"] = np.ascontiguousarray(inst.row_active, np.uint8)
        s.row_active = _b(keep["row_active"])
    s.obj_ub = float("inf")
    if inst.cut_col is not None and len(inst.cut_col):
        s.obj_ub = inst.cut_rhs + getattr(inst, "obj_const", 0.0)      # the raw incumbent value (fixObjBins_)
        keep["cut_col"] = np.ascontiguousarray(inst.cut_col, np.int32)
        keep["cut_val"] = np.ascontiguousarray(inst.cut_val, np.float64)
        s.cut_k, s.cut_col, s.cut_val, s.cut_rhs = len(keep["cut_col"]), _i(keep["cut_col"]), _d(keep["cut_val"]), inst.cut_rhs
    return s


def _nl_struct(t, keep, quad=None):
    """t: Tapes or None; quad: QuadCons or None (constraints with a QuadraticFunction, checked by chkRed_ only)."""
    s = OrcNl()
    conv = {np.int32: _i, np.uint8: _b, np.float64: _d}
    if t is None:
        s.n_cons = 0
        z = np.zeros(1, np.int32); keep["nl_zero"] = z
        s.tape_ptr = _i(z); s.lin_ptr = _i(z)
    else:
        s.n_cons = t.n_cons
        for name, typ in (("tape_ptr", np.int32), ("op", np.uint8), ("arg0", np.int32), ("arg1", np.int32),
                          ("cnst", np.float64), ("child", np.int32), ("lin_ptr", np.int32), ("lin_col", np.int32),
                          ("lin_val", np.float64), ("c_lb", np.float64), ("c_ub", np.float64)):
            a = np.ascontiguousarray(getattr(t, name), typ)
            keep["nl_" + name] = a
            setattr(s, name, conv[typ](a))
    if quad is None:
        quad = getattr(t, "quad", None)
    s.n_quad = 0
    if quad is not None and quad.n_quad > 0:
        s.n_quad = quad.n_quad
        for dst, src, typ in (("q_ptr", "q_ptr", np.int32), ("q_v1", "v1", np.int32), ("q_v2", "v2", np.int32),
                              ("q_coef", "coef", np.float64), ("q_lin_ptr", "lin_ptr", np.int32),
                              ("q_lin_col", "lin_col", np.int32), ("q_lin_val", "lin_val", np.float64),
                              ("q_lb", "q_lb", np.float64), ("q_ub", "q_ub", np.float64)):
            a = np.ascontiguousarray(getattr(quad, src), typ)
            keep["q_" + dst] = a
            setattr(s, dst, conv[typ](a))
    return s


class Oracle:
    """The plain-C restatement (oracle/fbbt_oracle.c)."""

    def __init__(self):
        build_oracle()
        self.lib = C.CDLL(ORACLE_SO)
        L = self.lib
        L.orc_lin_simple_presolve.argtypes = [C.POINTER(OrcLin), _dp, _dp, C.POINTER(OrcResult)]
        L.orc_lin_fixpoint_inplace.argtypes = [C.POINTER(OrcLin), _dp, _dp, C.POINTER(OrcResult)]
        L.orc_lin_fixpoint_jacobi.argtypes = [C.POINTER(OrcLin), _dp, _dp, C.c_int32, C.POINTER(OrcResult)]
        L.orc_lin_row_activity.argtypes = [C.POINTER(OrcLin), C.c_int32, _dp, _dp, _dp]
        L.orc_lin_jacobi_round_rows.argtypes = [C.POINTER(OrcLin), _dp, _dp, _dp, _dp]
        L.orc_lin_jacobi_round_rows.restype = C.c_int32
        L.orc_lin_jacobi_round_vars.argtypes = [C.POINTER(OrcLin), _dp, _dp, _dp, _dp, _ip]
        L.orc_lin_jacobi_round_vars.restype = C.c_int32
        L.orc_nl_compute_bounds.argtypes = [C.POINTER(OrcNl), C.c_int32, _dp, _dp, _dp, _dp]
        L.orc_nl_compute_bounds.restype = C.c_int32
        L.orc_nl_var_bound_mods.argtypes = [C.POINTER(OrcNl), C.c_int32, C.c_double, C.c_double, _dp, _dp, _ip]
        L.orc_nl_var_bound_mods.restype = C.c_int32
        L.orc_nl_simple_presolve.argtypes = [C.POINTER(OrcNl), _dp, _dp, C.POINTER(OrcResult)]
        L.orc_nl_simple_presolve_obj.argtypes = [C.POINTER(OrcNl), C.POINTER(OrcLin), _dp, _dp, C.POINTER(OrcResult)]
        L.orc_nl_chk_red.argtypes = [C.POINTER(OrcNl), _dp, _dp]
        L.orc_nl_chk_red.restype = C.c_int32
        L.orc_quad_compute_bounds.argtypes = [C.POINTER(OrcNl), C.c_int32, _dp, _dp, _dp, _dp]
        L.orc_quad_compute_bounds.restype = None
        L.orc_root_dup_rows.argtypes = [C.POINTER(OrcLin), _dp, _dp, _dp, _dp, C.c_int64, _ip, _ip, _bp]
        L.orc_root_dup_rows.restype = C.c_int64
        L.orc_root_redundant_rows.argtypes = [C.POINTER(OrcLin), _dp, _dp, _bp]
        L.orc_root_redundant_rows.restype = C.c_int64
        L.orc_root_coeff_imp.argtypes = [C.POINTER(OrcLin), _dp, _dp, C.c_int64, _ip, _ip, _dp, _ip, _dp]
        L.orc_root_coeff_imp.restype = C.c_int64
        L.orc_quad_simple_presolve.argtypes = [C.c_int32, _ip, _ip, C.c_int32, _ip, _ip, _ip, _bp, _dp, _dp, _lp]
        L.orc_quad_simple_presolve.restype = C.c_int64
        L.orc_quad_presolve_node.argtypes = [C.c_int32, _ip, _ip, C.c_int32, _ip, _ip, _ip, _bp, _dp, _dp, C.c_int32, _lp, _ip]
        L.orc_quad_presolve_node.restype = C.c_int32
        L.orc_nl_sweep.argtypes = [C.POINTER(OrcNl), _dp, _dp, C.POINTER(C.c_int64)]
        L.orc_nl_sweep.restype = C.c_int32
        L.orc_node_presolve.argtypes = [C.POINTER(OrcLin), C.POINTER(OrcNl), _dp, _dp, C.POINTER(OrcResult)]
        L.orc_batch_deltas.argtypes = [C.POINTER(OrcLin), C.POINTER(OrcNl), C.c_int32, C.c_int32, _dp, _dp, _lp, _ip, _bp,
                                       _dp, C.c_int32, _ip, _ip, _lp, C.c_int32, _ip, _ip, _bp, _dp]
        L.orc_batch_deltas.restype = C.c_double
        for f in ("orc_bounds_on_product",):
            getattr(L, f).argtypes = [C.c_int, C.c_double, C.c_double, C.c_double, C.c_double, _dp, _dp]
        L.orc_bounds_on_div.argtypes = [C.c_double] * 4 + [_dp, _dp]
        L.orc_bounds_on_recip.argtypes = [C.c_double] * 2 + [_dp, _dp]
        L.orc_bounds_on_square.argtypes = [C.c_double] * 2 + [_dp, _dp]

    # ---- linear ----
    def _run_lin(self, fn, inst, lb, ub, *extra):
        keep = {}
        s = _lin_struct(inst, keep)
        lb = np.array(lb, np.float64, copy=True); ub = np.array(ub, np.float64, copy=True)
        r = OrcResult()
        fn(C.byref(s), _d(lb), _d(ub), *extra, C.byref(r))
        return lb, ub, dict(verdict=r.verdict, rounds=r.rounds, nnz_updates=r.nnz_updates, n_mods=r.n_mods)

    def lin_simple_presolve(self, inst, lb=None, ub=None):
        return self._run_lin(self.lib.orc_lin_simple_presolve, inst, inst.lb if lb is None else lb,
                             inst.ub if ub is None else ub)

    def lin_fixpoint_inplace(self, inst, lb=None, ub=None):
        return self._run_lin(self.lib.orc_lin_fixpoint_inplace, inst, inst.lb if lb is None else lb,
                             inst.ub if ub is None else ub)

    def lin_fixpoint_jacobi(self, inst, lb=None, ub=None, max_rounds=0):
        return self._run_lin(self.lib.orc_lin_fixpoint_jacobi, inst, inst.lb if lb is None else lb,
                             inst.ub if ub is None else ub, C.c_int32(max_rounds))

    def lin_jacobi_round_rows(self, inst, lb, ub):
        keep = {}
        s = _lin_struct(inst, keep)
        lb = np.ascontiguousarray(lb, np.float64); ub = np.ascontiguousarray(ub, np.float64)
        nl = np.empty_like(lb); nu = np.empty_like(ub)
        inf = self.lib.orc_lin_jacobi_round_rows(C.byref(s), _d(lb), _d(ub), _d(nl), _d(nu))
        return nl, nu, int(inf)

    def lin_jacobi_round_vars(self, inst, lb, ub, nl, nu):
        keep = {}
        s = _lin_struct(inst, keep)
        lb = np.array(lb, np.float64, copy=True); ub = np.array(ub, np.float64, copy=True)
        nl = np.array(nl, np.float64, copy=True); nu = np.array(nu, np.float64, copy=True)
        ch = np.zeros(1, np.int32)
        inf = self.lib.orc_lin_jacobi_round_vars(C.byref(s), _d(lb), _d(ub), _d(nl), _d(nu), _i(ch))
        return lb, ub, int(inf), int(ch[0])

    def lin_row_activity(self, inst, row, lb, ub):
        keep = {}
        s = _lin_struct(inst, keep)
        out = np.zeros(4)
        lb = np.ascontiguousarray(lb, np.float64); ub = np.ascontiguousarray(ub, np.float64)
        self.lib.orc_lin_row_activity(C.byref(s), row, _d(lb), _d(ub), _d(out))
        return out

    # ---- root presolve row operations ----
    def root_dup_rows(self, inst, r1, r2, cap=1 << 16):
        """Duplicate-row candidates of LinearHandler::dupRows_: (h1, h2, pairs [k, 3] = (i, j, kind))."""
        keep = {}; s = _lin_struct(inst, keep)
        r1 = np.ascontiguousarray(r1, np.float64); r2 = np.ascontiguousarray(r2, np.float64)
        h1 = np.zeros(max(inst.m, 1)); h2 = np.zeros(max(inst.m, 1))
        pi = np.zeros(cap, np.int32); pj = np.zeros(cap, np.int32); pk = np.zeros(cap, np.uint8)
        k = int(self.lib.orc_root_dup_rows(C.byref(s), _d(r1), _d(r2), _d(h1), _d(h2), cap, _i(pi), _i(pj), _b(pk)))
        assert k <= cap
        return h1[:inst.m], h2[:inst.m], np.stack([pi[:k], pj[:k], pk[:k].astype(np.int32)], axis=1)

    def root_redundant_rows(self, inst, lb, ub):
        keep = {}; s = _lin_struct(inst, keep)
        lb = np.ascontiguousarray(lb, np.float64); ub = np.ascontiguousarray(ub, np.float64)
        out = np.zeros(max(inst.m, 1), np.uint8)
        self.lib.orc_root_redundant_rows(C.byref(s), _d(lb), _d(ub), _b(out))
        return out[:inst.m].astype(bool)

    def root_coeff_imp(self, inst, lb, ub, cap=1 << 20):
        """LinearHandler::coeffImp_ restated: (row, var, new coefficient, side, new row bound) arrays."""
        keep = {}; s = _lin_struct(inst, keep)
        lb = np.ascontiguousarray(lb, np.float64); ub = np.ascontiguousarray(ub, np.float64)
        row = np.zeros(cap, np.int32); var = np.zeros(cap, np.int32); coef = np.zeros(cap); side = np.zeros(cap, np.int32)
        bnd = np.zeros(cap)
        k = int(self.lib.orc_root_coeff_imp(C.byref(s), _d(lb), _d(ub), cap, _i(row), _i(var), _d(coef), _i(side), _d(bnd)))
        assert k <= cap
        return row[:k], var[:k], coef[:k], side[:k], bnd[:k]

    def quad_simple_presolve(self, rel, var_type, lb, ub):
        """QuadHandler::simplePresolve restated: one in-place sweep over the relations ``rel`` (instances.QuadRelations, in
        the handler's container order).  Returns (lb, ub, n_mods, n_inconsistent)."""
        lb = np.array(lb, np.float64); ub = np.array(ub, np.float64)
        vt = np.ascontiguousarray(var_type, np.uint8)
        bad = np.zeros(1, np.int64)
        a = [np.ascontiguousarray(x if len(x) else np.zeros(1), np.int32) for x in (rel.sq_x, rel.sq_y, rel.b_x0, rel.b_x1, rel.b_y)]
        k = int(self.lib.orc_quad_simple_presolve(len(rel.sq_x), _i(a[0]), _i(a[1]), len(rel.b_x0), _i(a[2]), _i(a[3]), _i(a[4]),
                                                  _b(vt), _d(lb), _d(ub), bad.ctypes.data_as(_lp)))
        return lb, ub, k, int(bad[0])

    def quad_presolve_node(self, rel, var_type, lb, ub, max_sweeps=0):
        """The propagation loop of QuadHandler::presolveNode restated (sweeps to the fixpoint, first inconsistency ends
        it).  Returns (lb, ub, infeasible, n_mods, n_sweeps)."""
        lb = np.array(lb, np.float64); ub = np.array(ub, np.float64)
        vt = np.ascontiguousarray(var_type, np.uint8)
        nm = np.zeros(1, np.int64); ns = np.zeros(1, np.int32)
        a = [np.ascontiguousarray(x if len(x) else np.zeros(1), np.int32) for x in (rel.sq_x, rel.sq_y, rel.b_x0, rel.b_x1, rel.b_y)]
        inf = int(self.lib.orc_quad_presolve_node(len(rel.sq_x), _i(a[0]), _i(a[1]), len(rel.b_x0), _i(a[2]), _i(a[3]), _i(a[4]),
                                                  _b(vt), _d(lb), _d(ub), max_sweeps, nm.ctypes.data_as(_lp), _i(ns)))
        return lb, ub, inf, int(nm[0]), int(ns[0])

    # ---- nonlinear ----
    def nl_compute_bounds(self, tapes, c, lb, ub):
        keep = {}; s = _nl_struct(tapes, keep)
        lb = np.ascontiguousarray(lb, np.float64); ub = np.ascontiguousarray(ub, np.float64)
        o = np.zeros(2)
        err = self.lib.orc_nl_compute_bounds(C.byref(s), c, _d(lb), _d(ub), _d(o[0:1]), _d(o[1:2]))
        return o[0], o[1], err

    def nl_var_bound_mods(self, tapes, c, lb_in, ub_in, lb, ub):
        keep = {}; s = _nl_struct(tapes, keep)
        lb = np.array(lb, np.float64, copy=True); ub = np.array(ub, np.float64, copy=True)
        nm = np.zeros(1, np.int32)
        st = self.lib.orc_nl_var_bound_mods(C.byref(s), c, lb_in, ub_in, _d(lb), _d(ub), _i(nm))
        return lb, ub, st, int(nm[0])

    def nl_chk_red(self, tapes, lb, ub, quad=None):
        """NlPresHandler::chkRed_ over the CGraph constraints and the QuadraticFunction constraints: 0 ok, 1 infeasible."""
        keep = {}; s = _nl_struct(tapes, keep, quad)
        lb = np.ascontiguousarray(lb, np.float64); ub = np.ascontiguousarray(ub, np.float64)
        return int(self.lib.orc_nl_chk_red(C.byref(s), _d(lb), _d(ub)))

    def quad_compute_bounds(self, quad, q, lb, ub):
        keep = {}; s = _nl_struct(None, keep, quad)
        lb = np.ascontiguousarray(lb, np.float64); ub = np.ascontiguousarray(ub, np.float64)
        o = np.zeros(2)
        self.lib.orc_quad_compute_bounds(C.byref(s), int(q), _d(lb), _d(ub), _d(o[0:1]), _d(o[1:2]))
        return o[0], o[1]

    def nl_simple_presolve(self, tapes, lb, ub, obj=None, quad=None):
        """obj: a LinearRows whose cut-off row is the (linear) objective, with an incumbent -> fixObjBins_ runs."""
        keep = {}; s = _nl_struct(tapes, keep, quad)
        lb = np.array(lb, np.float64, copy=True); ub = np.array(ub, np.float64, copy=True)
        r = OrcResult()
        if obj is not None:
            o = _lin_struct(obj, keep)
            self.lib.orc_nl_simple_presolve_obj(C.byref(s), C.byref(o), _d(lb), _d(ub), C.byref(r))
        else:
            self.lib.orc_nl_simple_presolve(C.byref(s), _d(lb), _d(ub), C.byref(r))
        return lb, ub, dict(verdict=r.verdict, rounds=r.rounds, n_mods=r.n_mods)

    def node_presolve(self, inst, tapes, lb, ub, quad=None):
        keep = {}
        s = _lin_struct(inst, keep) if inst is not None and inst.m > 0 else None
        g = _nl_struct(tapes, keep, quad) if (tapes is not None or quad is not None) else None
        lb = np.array(lb, np.float64, copy=True); ub = np.array(ub, np.float64, copy=True)
        r = OrcResult()
        self.lib.orc_node_presolve(C.byref(s) if s is not None else None, C.byref(g) if g is not None else None,
                                   _d(lb), _d(ub), C.byref(r))
        return lb, ub, dict(verdict=r.verdict, rounds=r.rounds, nnz_updates=r.nnz_updates, n_mods=r.n_mods)

    def batch_deltas(self, inst, tapes, mode, root_lb, root_ub, deltas, n_threads=1, mod_cap=0):
        """Node boxes given as branching deltas on a root box (the form of mntr_gpu_tighten_nodes), dealt to
        ``n_threads`` OpenMP threads.  mode 0 = lin_fixpoint_inplace, 1 = lin_simple_presolve, 2 = node_presolve.
        Returns dict(secs, verdict, rounds, nnz, mod_cnt, mod_var, mod_up, mod_val): the mods of box b are the
        first min(mod_cnt[b], mod_cap) entries of row b, ascending (variable, side)."""
        keep = {}
        s = _lin_struct(inst, keep)
        g = _nl_struct(tapes, keep) if tapes is not None else None
        ptr, var, up, val = _delta_arrays(deltas)
        nb = len(ptr) - 1
        rl = np.ascontiguousarray(root_lb, np.float64); ru = np.ascontiguousarray(root_ub, np.float64)
        v = np.zeros(nb, np.int32); r = np.zeros(nb, np.int32); z = np.zeros(nb, np.int64)
        cap = int(mod_cap)
        mc = np.zeros(nb, np.int32)
        mv = np.zeros((nb, max(cap, 1)), np.int32); mu = np.zeros((nb, max(cap, 1)), np.uint8); mx = np.zeros((nb, max(cap, 1)))
        secs = self.lib.orc_batch_deltas(C.byref(s), C.byref(g) if g is not None else None, int(mode), nb, _d(rl), _d(ru),
                                         _l(ptr), _i(var), _b(up), _d(val), int(n_threads), _i(v), _i(r), _l(z),
                                         cap, _i(mc) if cap > 0 else None, _i(mv), _b(mu), _d(mx))
        return dict(secs=secs, verdict=v, rounds=r, nnz=z, mod_cnt=mc, mod_var=mv, mod_up=mu, mod_val=mx)

    def bounds_on_product(self, z, l0, u0, l1, u1):
        o = np.zeros(2); self.lib.orc_bounds_on_product(int(z), l0, u0, l1, u1, _d(o[0:1]), _d(o[1:2])); return tuple(o)

    def bounds_on_div(self, l0, u0, l1, u1):
        o = np.zeros(2); self.lib.orc_bounds_on_div(l0, u0, l1, u1, _d(o[0:1]), _d(o[1:2])); return tuple(o)


class Reference:
    """A Minotaur::Problem built inside oracle/_ref/libminotaur_ref.so, driven by the
    reference's own LinearHandler / NlPresHandler / CGraph code."""

    _lib = None

    @classmethod
    def lib(cls):
        if cls._lib is None:
            L = C.CDLL(REF_SO)
            L.ref_create.restype = C.c_void_p
            L.ref_create.argtypes = [C.c_int32, C.c_int32, _ip, _ip, _dp, _dp, _dp, _bp, _dp, _dp]
            L.ref_add_nl.argtypes = [C.c_void_p, C.c_int32, _bp, _ip, _ip, _dp, _ip, C.c_int32, _ip, _dp,
                                     C.c_double, C.c_double]
            L.ref_add_nl.restype = C.c_int32
            L.ref_finish.argtypes = [C.c_void_p]
            L.ref_set_objective.argtypes = [C.c_void_p, C.c_int32, _ip, _dp, C.c_double]
            L.ref_set_incumbent.argtypes = [C.c_void_p, C.c_int32, C.c_double]
            L.ref_set_box.argtypes = [C.c_void_p, _dp, _dp]
            L.ref_get_box.argtypes = [C.c_void_p, _dp, _dp]
            L.ref_lin_fixpoint.argtypes = [C.c_void_p, _ip, C.POINTER(C.c_int64)]
            L.ref_lin_fixpoint_counted.argtypes = [C.c_void_p, C.c_int32, _ip, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
            L.ref_lin_simple_presolve.argtypes = [C.c_void_p, C.POINTER(C.c_int64)]
            L.ref_nl_simple_presolve.argtypes = [C.c_void_p, C.POINTER(C.c_int64)]
            L.ref_node_presolve.argtypes = [C.c_void_p, C.POINTER(C.c_int64)]
            L.ref_nl_compute_bounds.argtypes = [C.c_void_p, C.c_int32, _dp, _dp]
            L.ref_nl_var_bound_mods.argtypes = [C.c_void_p, C.c_int32, C.c_double, C.c_double, _ip]
            L.ref_nl_dq_ops.argtypes = [C.c_void_p, C.c_int32, C.c_int32, _ip]
            L.ref_row_activity.argtypes = [C.c_void_p, C.c_int32, _dp]
            L.ref_time_boxes.argtypes = [C.c_void_p, C.c_int32, C.c_int32, _dp, _dp, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
            L.ref_time_boxes.restype = C.c_double
            L.ref_time_deltas.argtypes = [C.c_void_p, C.c_int32, C.c_int32, _dp, _dp, _lp, _ip, _bp, _dp,
                                          C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
            L.ref_time_deltas.restype = C.c_double
            L.ref_add_nl_batch.argtypes = [C.c_void_p, C.c_int32, _ip, _bp, _ip, _ip, _dp, _ip, _ip, _ip, _dp, _dp, _dp]
            L.ref_add_nl_batch.restype = C.c_int32
            L.ref_add_quad.argtypes = [C.c_void_p, C.c_int32, _ip, _ip, _dp, C.c_int32, _ip, _dp, C.c_double, C.c_double]
            L.ref_add_quad.restype = C.c_int32
            L.ref_nl_chk_red.argtypes = [C.c_void_p]
            L.ref_nl_chk_red.restype = C.c_int32
            L.ref_quad_compute_bounds.argtypes = [C.c_void_p, C.c_int32, _dp, _dp]
            L.ref_draw_dup_vectors.argtypes = [C.c_void_p, C.c_uint32, _dp, _dp]
            L.ref_dup_rows.argtypes = [C.c_void_p, C.c_uint32, _bp, _dp, _dp]
            L.ref_dup_rows_replay.argtypes = [C.c_void_p, C.c_int64, _ip, _ip, _bp, _dp, _bp, _dp, _dp]
            L.ref_redundant_rows.argtypes = [C.c_void_p, _bp]
            L.ref_coeff_imp.restype = C.c_int64
            L.ref_quad_simple_presolve.restype = C.c_int64
            L.ref_quad_simple_presolve.argtypes = [C.c_int32, _bp, _dp, _dp, C.c_int32, _ip, _ip, C.c_int32, _ip, _ip, _ip]
            L.ref_coeff_imp.argtypes = [C.c_void_p, C.c_int64, _ip, _ip, _dp, _ip, _dp]
            L.ref_quad_presolve_node.argtypes = [C.c_int32, _bp, _dp, _dp, C.c_int32, _ip, _ip, C.c_int32, _ip, _ip, _ip, C.c_int32,
                                                 _dp, _dp, _ip, _lp]
            L.ref_destroy.argtypes = [C.c_void_p]
            cls._lib = L
        return cls._lib

    def __init__(self, inst, tapes=None, quad=None):
        L = self.lib()
        self.n = inst.n
        rp = np.ascontiguousarray(inst.row_ptr, np.int32); col = np.ascontiguousarray(inst.col, np.int32)
        val = np.ascontiguousarray(inst.val, np.float64)
        rl = np.ascontiguousarray(inst.row_lb, np.float64); ru = np.ascontiguousarray(inst.row_ub, np.float64)
        vt = np.ascontiguousarray(inst.var_type, np.uint8)
        lb = np.ascontiguousarray(inst.lb, np.float64); ub = np.ascontiguousarray(inst.ub, np.float64)
        if inst.row_active is not None and not np.all(inst.row_active):
            raise ValueError("Reference harness: deleted rows are not supported; drop them from the CSR")
        self.h = L.ref_create(inst.m, inst.n, _i(rp), _i(col), _d(val), _d(rl), _d(ru), _b(vt), _d(lb), _d(ub))
        if tapes is not None:
            keep = {}
            g = _nl_struct(tapes, keep)
            L.ref_add_nl_batch(self.h, tapes.n_cons, g.tape_ptr, g.op, g.arg0, g.arg1, g.cnst, g.child, g.lin_ptr,
                               g.lin_col, g.lin_val, g.c_lb, g.c_ub)
        if quad is not None:
            for q in range(quad.n_quad):
                b, e = int(quad.q_ptr[q]), int(quad.q_ptr[q + 1])
                lb_, le_ = int(quad.lin_ptr[q]), int(quad.lin_ptr[q + 1])
                a1 = np.ascontiguousarray(quad.v1[b:e], np.int32); a2 = np.ascontiguousarray(quad.v2[b:e], np.int32)
                cf = np.ascontiguousarray(quad.coef[b:e], np.float64)
                lc = np.ascontiguousarray(quad.lin_col[lb_:le_] if le_ > lb_ else np.zeros(1, np.int32), np.int32)
                lv = np.ascontiguousarray(quad.lin_val[lb_:le_] if le_ > lb_ else np.zeros(1), np.float64)
                L.ref_add_quad(self.h, e - b, _i(a1), _i(a2), _d(cf), le_ - lb_, _i(lc), _d(lv), float(quad.q_lb[q]),
                               float(quad.q_ub[q]))
        if inst.cut_col is not None and len(inst.cut_col):
            # the cut-off row c.x <= cut_rhs as the reference meets it: a linear objective c.x + obj_const and an
            # incumbent of value cut_rhs + obj_const in the solution pool (LinearHandler.cpp:1636-1640)
            cc = np.ascontiguousarray(inst.cut_col, np.int32); cv = np.ascontiguousarray(inst.cut_val, np.float64)
            const = float(getattr(inst, "obj_const", 0.0))
            L.ref_set_objective(self.h, len(cc), _i(cc), _d(cv), const)
            L.ref_set_incumbent(self.h, 1, float(inst.cut_rhs) + const)
        L.ref_finish(self.h)

    def set_incumbent(self, value=None):
        """value None removes the incumbent (no cut-off row)."""
        self.lib().ref_set_incumbent(self.h, 0 if value is None else 1, 0.0 if value is None else float(value))

    def close(self):
        if self.h:
            self.lib().ref_destroy(self.h); self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_box(self, lb, ub):
        lb = np.ascontiguousarray(lb, np.float64); ub = np.ascontiguousarray(ub, np.float64)
        self.lib().ref_set_box(self.h, _d(lb), _d(ub))

    def get_box(self) -> Tuple[np.ndarray, np.ndarray]:
        lb = np.zeros(self.n); ub = np.zeros(self.n)
        self.lib().ref_get_box(self.h, _d(lb), _d(ub))
        return lb, ub

    def lin_fixpoint(self, lb, ub, counted=False, max_rounds=0):
        self.set_box(lb, ub)
        rounds = C.c_int32(0); nm = C.c_int64(0); nnz = C.c_int64(0)
        if counted:
            v = self.lib().ref_lin_fixpoint_counted(self.h, max_rounds, C.byref(rounds), C.byref(nm), C.byref(nnz))
        else:
            v = self.lib().ref_lin_fixpoint(self.h, C.byref(rounds), C.byref(nm))
        l, u = self.get_box()
        return l, u, dict(verdict=int(v), rounds=rounds.value, n_mods=nm.value, nnz_updates=nnz.value)

    def lin_simple_presolve(self, lb, ub):
        self.set_box(lb, ub)
        nm = C.c_int64(0)
        v = self.lib().ref_lin_simple_presolve(self.h, C.byref(nm))
        l, u = self.get_box()
        return l, u, dict(verdict=int(v), n_mods=nm.value)

    def nl_simple_presolve(self, lb, ub):
        self.set_box(lb, ub)
        nm = C.c_int64(0)
        v = self.lib().ref_nl_simple_presolve(self.h, C.byref(nm))
        l, u = self.get_box()
        return l, u, dict(verdict=int(v), n_mods=nm.value)

    def node_presolve(self, lb, ub):
        self.set_box(lb, ub)
        nm = C.c_int64(0)
        v = self.lib().ref_node_presolve(self.h, C.byref(nm))
        l, u = self.get_box()
        return l, u, dict(verdict=int(v), n_mods=nm.value)

    def nl_compute_bounds(self, c, lb, ub):
        self.set_box(lb, ub)
        o = np.zeros(2)
        err = self.lib().ref_nl_compute_bounds(self.h, c, _d(o[0:1]), _d(o[1:2]))
        return o[0], o[1], err

    def nl_var_bound_mods(self, c, lb_in, ub_in, lb, ub):
        self.set_box(lb, ub)
        nm = np.zeros(1, np.int32)
        st = self.lib().ref_nl_var_bound_mods(self.h, c, lb_in, ub_in, _i(nm))
        l, u = self.get_box()
        return l, u, int(st), int(nm[0])

    def draw_dup_vectors(self, seed):
        """The two random vectors LinearHandler::dupRows_ draws after srand(seed)."""
        r1 = np.zeros(self.n); r2 = np.zeros(self.n)
        self.lib().ref_draw_dup_vectors(self.h, int(seed), _d(r1), _d(r2))
        return r1, r2

    def dup_rows(self, seed, m):
        """The reference's own dupRows_ (generator seeded): (deleted [m], row_lb, row_ub).  Modifies the problem."""
        d = np.zeros(max(m, 1), np.uint8); rl = np.zeros(max(m, 1)); ru = np.zeros(max(m, 1))
        self.lib().ref_dup_rows(self.h, int(seed), _b(d), _d(rl), _d(ru))
        return d[:m].astype(bool), rl[:m], ru[:m]

    def dup_rows_replay(self, pairs, h1, m):
        """dupRows_'s loop driven by an external candidate list (treatDupRows_ is the reference's).  Modifies the problem."""
        pi = np.ascontiguousarray(pairs[:, 0], np.int32); pj = np.ascontiguousarray(pairs[:, 1], np.int32)
        pk = np.ascontiguousarray(pairs[:, 2], np.uint8); h1 = np.ascontiguousarray(h1, np.float64)
        if len(pi) == 0:
            pi = np.zeros(1, np.int32); pj = np.zeros(1, np.int32); pk = np.zeros(1, np.uint8)
        d = np.zeros(max(m, 1), np.uint8); rl = np.zeros(max(m, 1)); ru = np.zeros(max(m, 1))
        self.lib().ref_dup_rows_replay(self.h, len(pairs), _i(pi), _i(pj), _b(pk), _d(h1), _b(d), _d(rl), _d(ru))
        return d[:m].astype(bool), rl[:m], ru[:m]

    @classmethod
    def quad_simple_presolve(cls, rel, var_type, lb, ub):
        """The reference's own QuadHandler::simplePresolve on a fresh problem holding the relations.  (lb, ub, n_mods)"""
        lb = np.array(lb, np.float64); ub = np.array(ub, np.float64)
        vt = np.ascontiguousarray(var_type, np.uint8)
        a = [np.ascontiguousarray(x if len(x) else np.zeros(1), np.int32) for x in (rel.sq_x, rel.sq_y, rel.b_x0, rel.b_x1, rel.b_y)]
        k = int(cls.lib().ref_quad_simple_presolve(len(lb), _b(vt), _d(lb), _d(ub), len(rel.sq_x), _i(a[0]), _i(a[1]),
                                                   len(rel.b_x0), _i(a[2]), _i(a[3]), _i(a[4])))
        return lb, ub, k

    @classmethod
    def quad_presolve_node(cls, rel, var_type, root_lb, root_ub, L, U):
        """The reference's own QuadHandler::presolveNode (as at every node after the first: no tightenQuad_) on the boxes
        L, U [n_boxes][n] of a problem built on the root box.  Returns (lb, ub, verdict, n_mods)."""
        L = np.array(L, np.float64, ndmin=2); U = np.array(U, np.float64, ndmin=2)
        nb, n = L.shape
        vt = np.ascontiguousarray(var_type, np.uint8)
        rl = np.ascontiguousarray(root_lb, np.float64); ru = np.ascontiguousarray(root_ub, np.float64)
        a = [np.ascontiguousarray(x if len(x) else np.zeros(1), np.int32) for x in (rel.sq_x, rel.sq_y, rel.b_x0, rel.b_x1, rel.b_y)]
        verdict = np.zeros(nb, np.int32); nm = np.zeros(nb, np.int64)
        cls.lib().ref_quad_presolve_node(n, _b(vt), _d(rl), _d(ru), len(rel.sq_x), _i(a[0]), _i(a[1]), len(rel.b_x0), _i(a[2]),
                                         _i(a[3]), _i(a[4]), nb, _d(L), _d(U), _i(verdict), nm.ctypes.data_as(_lp))
        return L, U, verdict, nm

    def coeff_imp(self, lb, ub, cap=1 << 20):
        """The reference's own LinearHandler::coeffImp_ on the box.  Modifies the problem (use a fresh Reference).
        Returns (row, var, new coefficient, side, new row bound) arrays."""
        self.set_box(lb, ub)
        row = np.zeros(cap, np.int32); var = np.zeros(cap, np.int32); coef = np.zeros(cap); side = np.zeros(cap, np.int32)
        bnd = np.zeros(cap)
        k = int(self.lib().ref_coeff_imp(self.h, cap, _i(row), _i(var), _d(coef), _i(side), _d(bnd)))
        assert k <= cap
        return row[:k], var[:k], coef[:k], side[:k], bnd[:k]

    def redundant_rows(self, lb, ub, m):
        self.set_box(lb, ub)
        out = np.zeros(max(m, 1), np.uint8)
        self.lib().ref_redundant_rows(self.h, _b(out))
        return out[:m].astype(bool)

    def nl_chk_red(self, lb, ub):
        """NlPresHandler::chkRed_ alone on the box: 1 = infeasible."""
        self.set_box(lb, ub)
        return int(self.lib().ref_nl_chk_red(self.h))

    def quad_compute_bounds(self, q, lb, ub):
        self.set_box(lb, ub)
        o = np.zeros(2)
        self.lib().ref_quad_compute_bounds(self.h, int(q), _d(o[0:1]), _d(o[1:2]))
        return o[0], o[1]

    def nl_dq_ops(self, c):
        ops = np.zeros(4096, np.int32)
        k = self.lib().ref_nl_dq_ops(self.h, c, 4096, _i(ops))
        return ops[:k].copy()

    def row_activity(self, row, lb, ub):
        self.set_box(lb, ub)
        out = np.zeros(4)
        self.lib().ref_row_activity(self.h, row, _d(out))
        return out

    def time_boxes(self, mode, lbs, ubs):
        lbs = np.ascontiguousarray(lbs, np.float64); ubs = np.ascontiguousarray(ubs, np.float64)
        nb = lbs.shape[0] if lbs.ndim == 2 else 1
        nnz = C.c_int64(0); ninf = C.c_int64(0)
        secs = self.lib().ref_time_boxes(self.h, mode, nb, _d(lbs), _d(ubs), C.byref(nnz), C.byref(ninf))
        return secs, nnz.value, ninf.value


def _ref_time_deltas(self, mode, root_lb, root_ub, deltas):
    """ref_time_boxes for boxes given as branching deltas on a root box; returns (secs, nnz, n_infeasible)."""
    ptr, var, up, val = _delta_arrays(deltas)
    rl = np.ascontiguousarray(root_lb, np.float64); ru = np.ascontiguousarray(root_ub, np.float64)
    nnz = C.c_int64(0); ninf = C.c_int64(0)
    secs = self.lib().ref_time_deltas(self.h, int(mode), len(ptr) - 1, _d(rl), _d(ru), _l(ptr), _i(var), _b(up), _d(val),
                                      C.byref(nnz), C.byref(ninf))
    return secs, nnz.value, ninf.value


Reference.time_deltas = _ref_time_deltas


def ref_read_mps(path: str) -> dict:
    """The reference's own Reader::readMps (Reader.cpp:42-473) on `path`, dumped flat (fixture generator of
    minotaur_b200/mps_reader.py)."""
    L = Reference.lib()
    _u8 = C.POINTER(C.c_uint8)
    L.ref_read_mps.argtypes = [C.c_char_p, _ip, _ip, _ip, _ip, _ip, _dp, _dp, _dp, _u8, _dp, _dp, _ip, _ip, _dp, _dp]
    L.ref_read_mps.restype = C.c_int32
    m, n, nnz, ok = C.c_int32(0), C.c_int32(0), C.c_int32(0), C.c_int32(0)
    oc = C.c_double(0.0)
    null_i, null_d, null_u = _ip(), _dp(), _u8()
    err = L.ref_read_mps(path.encode(), C.byref(m), C.byref(n), C.byref(nnz), null_i, null_i, null_d, null_d, null_d,
                         null_u, null_d, null_d, C.byref(ok), null_i, null_d, C.byref(oc))
    if err != 0:
        raise RuntimeError(f"Reader::readMps failed with code {err}")
    out = dict(row_ptr=np.zeros(m.value + 1, np.int32), col=np.zeros(nnz.value, np.int32), val=np.zeros(nnz.value),
               row_lb=np.zeros(m.value), row_ub=np.zeros(m.value), var_type=np.zeros(n.value, np.uint8),
               lb=np.zeros(n.value), ub=np.zeros(n.value), obj_col=np.zeros(ok.value, np.int32), obj_val=np.zeros(ok.value))
    def ip(a): return a.ctypes.data_as(_ip)
    def dp(a): return a.ctypes.data_as(_dp)
    L.ref_read_mps(path.encode(), C.byref(m), C.byref(n), C.byref(nnz), ip(out["row_ptr"]), ip(out["col"]), dp(out["val"]),
                   dp(out["row_lb"]), dp(out["row_ub"]), out["var_type"].ctypes.data_as(_u8), dp(out["lb"]), dp(out["ub"]),
                   C.byref(ok), ip(out["obj_col"]), dp(out["obj_val"]), C.byref(oc))
    out["obj_const"] = np.array([oc.value])
    return out
