"""Generates tests/golden/linear_cases.npz and nl_cases.npz from THE REFERENCE'S OWN CODE.

Run in the build container (needs oracle/_ref/libminotaur_ref.so, which oracle/Makefile builds
from /root/reference/src/base/*.cpp):

    python tests/golden/make_golden.py

Every output array below was produced by LinearHandler / NlPresHandler / CGraph of the
reference, driven through oracle/ref_harness.cpp:
  raw_*   LinearHandler::simplePresolve            (LinearHandler.cpp:1605-1653)
  fix_*   the status-honouring fixpoint driver     
  act     getLfBnds_ / getSingLfBnds_ of every row (LinearHandler.cpp:1237-1319)
  nl_*    CGraph::computeBounds / varBoundMods / NlPresHandler::simplePresolve
The fixtures are the committed pin of oracle/fbbt_oracle.c for boxes where the GPU box has no
/root/reference.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from minotaur_b200.instances import (Expr, OpAbs, OpCeil, OpFloor, OpLog10, attach_binary_objective, attach_cutoff,  # noqa: E402
                                     branch_boxes, build_tapes,
                                     make_knapsack_setcover, make_minlp, make_sparse_milp, LinearRows, INF)
from oracle.pyoracle import Reference  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def linear_cases():
    out = {}
    specs = [
        ("intdata", dict(m=60, n=60, nnz_per_row=5, seed=11, real_data=False)),
        ("realdata", dict(m=60, n=50, nnz_per_row=6, seed=12, real_data=True)),
        ("inf", dict(m=70, n=60, nnz_per_row=5, seed=13, real_data=True, inf_frac=(0.15, 0.1, 0.05))),
        ("wide", dict(m=40, n=120, nnz_per_row=9, seed=14, real_data=False, inf_frac=(0.05, 0.05, 0.0))),
    ]
    names = []
    for name, kw in specs:
        inst = make_sparse_milp(**kw)
        names.append(name)
        _emit_linear(out, name, inst, n_boxes=10, seed=kw["seed"])
    inst = make_knapsack_setcover(m=80, n=60, nnz_per_row=6, seed=15)
    names.append("knap")
    _emit_linear(out, "knap", inst, n_boxes=10, seed=15)
    # objective cut-off row (LinearHandler::varBndsFromObj_): a linear objective plus an incumbent in the pool
    for name, kw, k, slack in (("cut_int", dict(m=60, n=60, nnz_per_row=5, seed=16, real_data=False), 25, 30.0),
                               ("cut_real", dict(m=70, n=60, nnz_per_row=6, seed=17, real_data=True), 40, 30.0)):
        inst = make_sparse_milp(**kw)
        ref = Reference(inst)
        tl, tu, _ = ref.lin_fixpoint(inst.lb, inst.ub)
        ref.close()
        inst = attach_cutoff(inst, k, kw["seed"], slack, box=(tl, tu))
        names.append(name)
        _emit_linear(out, name, inst, n_boxes=10, seed=kw["seed"])
    out["names"] = np.array(names)
    return out


def _emit_linear(out, name, inst, n_boxes, seed):
    ref = Reference(inst)
    lbs, ubs = branch_boxes(inst.lb, inst.ub, inst.var_type, n_boxes, seed=seed, max_depth=8)
    lbs[0], ubs[0] = inst.lb, inst.ub          # box 0 = the root box
    for k in ("row_ptr", "col", "val", "row_lb", "row_ub", "var_type"):
        out[f"{name}.{k}"] = getattr(inst, k)
    out[f"{name}.shape"] = np.array([inst.m, inst.n])
    if inst.cut_col is not None:
        out[f"{name}.cut_col"], out[f"{name}.cut_val"] = inst.cut_col, inst.cut_val
        out[f"{name}.cut_rhs"] = np.array([inst.cut_rhs])
    out[f"{name}.lbs"], out[f"{name}.ubs"] = lbs, ubs
    raw_lb, raw_ub, raw_v, raw_nm = [], [], [], []
    fix_lb, fix_ub, fix_v, fix_r, fix_nnz = [], [], [], [], []
    act = []
    for b in range(n_boxes):
        l, u, r = ref.lin_simple_presolve(lbs[b], ubs[b])
        raw_lb.append(l); raw_ub.append(u); raw_v.append(r["verdict"]); raw_nm.append(r["n_mods"])
        l, u, r = ref.lin_fixpoint(lbs[b], ubs[b], counted=True)
        l2, u2, r2 = ref.lin_fixpoint(lbs[b], ubs[b], counted=False)
        assert r["verdict"] == r2["verdict"] and (r["verdict"] or (np.array_equal(l, l2) and np.array_equal(u, u2)))
        fix_lb.append(l); fix_ub.append(u); fix_v.append(r["verdict"]); fix_r.append(r["rounds"])
        fix_nnz.append(r["nnz_updates"])
        act.append(np.stack([ref.row_activity(i, lbs[b], ubs[b]) for i in range(inst.m)]))
    out[f"{name}.raw_lb"], out[f"{name}.raw_ub"] = np.array(raw_lb), np.array(raw_ub)
    out[f"{name}.raw_verdict"], out[f"{name}.raw_nmods"] = np.array(raw_v, np.int32), np.array(raw_nm, np.int64)
    out[f"{name}.fix_lb"], out[f"{name}.fix_ub"] = np.array(fix_lb), np.array(fix_ub)
    out[f"{name}.fix_verdict"], out[f"{name}.fix_rounds"] = np.array(fix_v, np.int32), np.array(fix_r, np.int32)
    out[f"{name}.fix_nnz"] = np.array(fix_nnz, np.int64)
    out[f"{name}.act"] = np.array(act)
    ref.close()


def nl_expr_cases():
    """Hand-written CGraph constraints covering every implemented opcode and the reference's quirks
    (SURVEY.md section 7 hard part 3).  Each: (expr, linear part, c_lb, c_ub), on 6 variables."""
    v = Expr.v
    return [
        (v(0) * v(1), [], 1.0, 4.0),                                   # 1 <= x*y <= 4
        (v(0) * v(1), [(2, 2.0)], -INF, 6.0),                          # x*y + 2z <= 6
        (v(0).sqr(), [], -INF, 4.0),                                   # reverse OpSqr is a no-op
        (v(3).sqrt(), [], 2.0, 3.0),                                   # reverse OpSqrt: child >= 0 only
        (v(0).powk(4), [], -INF, 16.0),                                # even PowK reverse works
        (Expr.sumlist([v(0).sqr(), v(1).sqr()]), [], -INF, 9.0),
        (Expr.sumlist([v(0), v(1), v(2)]), [], -INF, 3.0),
        (v(0) + v(1), [], 1.0, 2.0),
        (v(0) - v(1), [], -1.0, 0.5),
        (v(0) / v(3), [], 0.5, 2.0),
        (-(v(0) * v(2)), [(1, 1.0)], -2.0, 2.0),
        (v(4).exp(), [], -INF, 5.0),
        (v(3).log(), [], 0.0, 1.0),
        (v(5).abs(), [], -INF, 1.5),
        ((v(0) + 3.0) * v(1), [], -INF, 8.0),
        (Expr.unary(OpFloor, v(2)), [], 1.0, 2.0),
        (Expr.unary(OpCeil, v(2)), [], 1.0, 2.0),
        (Expr.unary(OpLog10, v(3)), [], 0.0, 1.0),
        ((v(0) * v(1)) + (v(1) * v(2)), [(3, -1.0)], -INF, 5.0),
        (Expr.sumlist([v(0) * v(1), v(2).sqr(), v(4)]), [(5, 0.5)], -3.0, 12.0),
    ]


def nl_cases():
    out = {}
    cons = nl_expr_cases()
    tapes = build_tapes(cons)
    n = 6
    for k in ("tape_ptr", "op", "arg0", "arg1", "cnst", "child", "lin_ptr", "lin_col", "lin_val", "c_lb", "c_ub"):
        out[f"expr.{k}"] = getattr(tapes, k)
    rng = np.random.default_rng(5)
    boxes_l, boxes_u = [], []
    base_l = np.array([-3.0, -2.0, 0.0, 0.5, -1.0, -4.0]); base_u = np.array([5.0, 6.0, 4.0, 30.0, 3.0, 4.0])
    boxes_l.append(base_l); boxes_u.append(base_u)
    boxes_l.append(np.array([0.5, 0.25, 0.0, 1.0, 0.0, -1.0])); boxes_u.append(np.array([10.0, 12.0, 5.0, 100.0, 2.0, 1.0]))
    boxes_l.append(np.array([-10.0, -10.0, -10.0, 0.0, -5.0, -3.0])); boxes_u.append(np.array([10.0, 10.0, 10.0, 100.0, 5.0, 3.0]))
    for _ in range(9):
        l = base_l + rng.random(n) * 2 - 1; u = l + rng.random(n) * 8 + 0.1
        l[3] = abs(l[3]) + 0.01; u[3] = l[3] + 10 * rng.random() + 0.1
        boxes_l.append(l); boxes_u.append(u)
    boxes_l, boxes_u = np.array(boxes_l), np.array(boxes_u)
    out["expr.lbs"], out["expr.ubs"] = boxes_l, boxes_u
    dummy = LinearRows(m=0, n=n, row_ptr=np.zeros(1, np.int32), col=np.zeros(0, np.int32), val=np.zeros(0),
                       row_lb=np.zeros(0), row_ub=np.zeros(0), var_type=np.array([4, 4, 1, 4, 4, 4], np.uint8),
                       lb=base_l, ub=base_u)
    out["expr.var_type"] = dummy.var_type
    nb, nc = boxes_l.shape[0], tapes.n_cons
    cb = np.zeros((nb, nc, 3)); vm_l = np.zeros((nb, nc, n)); vm_u = np.zeros((nb, nc, n)); vm_s = np.zeros((nb, nc, 2), np.int32)
    dq = []
    for c in range(nc):
        ref = Reference(dummy, build_tapes([cons[c]]))      # a fresh graph per constraint
        dq.append(ref.nl_dq_ops(0))
        for b in range(nb):
            lo, hi, err = ref.nl_compute_bounds(0, boxes_l[b], boxes_u[b])
            cb[b, c] = (lo, hi, err)
        ref.close()
        for b in range(nb):
            ref = Reference(dummy, build_tapes([cons[c]]))  # fresh: constant nodes keep state in the reference
            lf = 0.0
            l, u, st, nm = ref.nl_var_bound_mods(0, float(cons[c][2]), float(cons[c][3]), boxes_l[b], boxes_u[b])
            vm_l[b, c], vm_u[b, c], vm_s[b, c] = l, u, (st, nm)
            ref.close()
    out["expr.compute_bounds"], out["expr.vbm_lb"], out["expr.vbm_ub"], out["expr.vbm_status"] = cb, vm_l, vm_u, vm_s
    out["expr.dq_ops"] = np.array([np.pad(d, (0, 16 - len(d)), constant_values=-1) for d in dq], np.int32)

    # whole-handler runs on a small C5-shaped MINLP: NlPresHandler::simplePresolve and the node presolve
    lin, tp = make_minlp(n=40, n_cons=60, m_lin=20, seed=7)
    for k in ("row_ptr", "col", "val", "row_lb", "row_ub", "var_type"):
        out[f"minlp.{k}"] = getattr(lin, k)
    out["minlp.shape"] = np.array([lin.m, lin.n])
    for k in ("tape_ptr", "op", "arg0", "arg1", "cnst", "child", "lin_ptr", "lin_col", "lin_val", "c_lb", "c_ub"):
        out[f"minlp.t.{k}"] = getattr(tp, k)
    lbs, ubs = branch_boxes(lin.lb, lin.ub, lin.var_type, 12, seed=3, max_depth=6, continuous_too=True)
    lbs[0], ubs[0] = lin.lb, lin.ub
    out["minlp.lbs"], out["minlp.ubs"] = lbs, ubs
    nl_l, nl_u, nl_v, nd_l, nd_u, nd_v = [], [], [], [], [], []
    for b in range(lbs.shape[0]):
        ref = Reference(lin, tp)
        l, u, r = ref.nl_simple_presolve(lbs[b], ubs[b])
        nl_l.append(l); nl_u.append(u); nl_v.append(r["verdict"])
        ref.close()
        ref = Reference(lin, tp)
        l, u, r = ref.node_presolve(lbs[b], ubs[b])
        nd_l.append(l); nd_u.append(u); nd_v.append(r["verdict"])
        ref.close()
    out["minlp.nl_lb"], out["minlp.nl_ub"], out["minlp.nl_verdict"] = np.array(nl_l), np.array(nl_u), np.array(nl_v, np.int32)
    out["minlp.node_lb"], out["minlp.node_ub"], out["minlp.node_verdict"] = np.array(nd_l), np.array(nd_u), np.array(nd_v, np.int32)

    # the same two runs with a linear objective over binaries and an incumbent: LinearHandler adds the cut-off row
    # (varBndsFromObj_), NlPresHandler adds fixObjBins_ (which compares against the raw incumbent value)
    lin, tp = make_minlp(n=300, n_cons=200, m_lin=60, seed=35)
    lin = attach_binary_objective(lin, n_bin=14, n_other=2, seed=35, slack=3.0, const=2.25, coef_hi=40)
    for k in ("row_ptr", "col", "val", "row_lb", "row_ub", "var_type", "cut_col", "cut_val"):
        out[f"minlp_obj.{k}"] = getattr(lin, k)
    out["minlp_obj.shape"] = np.array([lin.m, lin.n])
    out["minlp_obj.cut_rhs_const"] = np.array([lin.cut_rhs, lin.obj_const])
    for k in ("tape_ptr", "op", "arg0", "arg1", "cnst", "child", "lin_ptr", "lin_col", "lin_val", "c_lb", "c_ub"):
        out[f"minlp_obj.t.{k}"] = getattr(tp, k)
    lbs, ubs = branch_boxes(lin.lb, lin.ub, lin.var_type, 10, seed=35, max_depth=5, continuous_too=True)
    lbs[0], ubs[0] = lin.lb, lin.ub
    out["minlp_obj.lbs"], out["minlp_obj.ubs"] = lbs, ubs
    nl_l, nl_u, nl_v, nd_l, nd_u, nd_v = [], [], [], [], [], []
    for b in range(lbs.shape[0]):
        ref = Reference(lin, tp)
        l, u, r = ref.nl_simple_presolve(lbs[b], ubs[b])
        nl_l.append(l); nl_u.append(u); nl_v.append(r["verdict"])
        ref.close()
        ref = Reference(lin, tp)
        l, u, r = ref.node_presolve(lbs[b], ubs[b])
        nd_l.append(l); nd_u.append(u); nd_v.append(r["verdict"])
        ref.close()
    out["minlp_obj.nl_lb"], out["minlp_obj.nl_ub"], out["minlp_obj.nl_verdict"] = np.array(nl_l), np.array(nl_u), np.array(nl_v, np.int32)
    out["minlp_obj.node_lb"], out["minlp_obj.node_ub"], out["minlp_obj.node_verdict"] = np.array(nd_l), np.array(nd_u), np.array(nd_v, np.int32)
    return out


def _emit_minlp(out, name, lin, tp, lbs, ubs):
    """NlPresHandler::simplePresolve, LinearHandler::simplePresolve, the linear fixpoint driver and the node
    presolve (LinearHandler then NlPresHandler) of the reference on every box."""
    for k in ("row_ptr", "col", "val", "row_lb", "row_ub", "var_type"):
        out[f"{name}.{k}"] = getattr(lin, k)
    out[f"{name}.shape"] = np.array([lin.m, lin.n])
    if lin.cut_col is not None:
        out[f"{name}.cut_col"], out[f"{name}.cut_val"] = lin.cut_col, lin.cut_val
        out[f"{name}.cut_rhs_const"] = np.array([lin.cut_rhs, lin.obj_const])
    for k in ("tape_ptr", "op", "arg0", "arg1", "cnst", "child", "lin_ptr", "lin_col", "lin_val", "c_lb", "c_ub"):
        out[f"{name}.t.{k}"] = getattr(tp, k)
    out[f"{name}.lbs"], out[f"{name}.ubs"] = lbs, ubs
    acc = {k: [] for k in ("nl_lb", "nl_ub", "nl_verdict", "node_lb", "node_ub", "node_verdict", "raw_lb", "raw_ub",
                           "raw_verdict", "fix_lb", "fix_ub", "fix_verdict", "fix_rounds", "fix_nnz")}
    for b in range(lbs.shape[0]):
        for key, call in (("nl", "nl_simple_presolve"), ("node", "node_presolve"), ("raw", "lin_simple_presolve")):
            ref = Reference(lin, tp)            # fresh graphs: constant nodes keep state in the reference
            l, u, r = getattr(ref, call)(lbs[b], ubs[b])
            acc[f"{key}_lb"].append(l); acc[f"{key}_ub"].append(u); acc[f"{key}_verdict"].append(r["verdict"])
            ref.close()
        ref = Reference(lin, tp)
        l, u, r = ref.lin_fixpoint(lbs[b], ubs[b], counted=True)
        acc["fix_lb"].append(l); acc["fix_ub"].append(u); acc["fix_verdict"].append(r["verdict"])
        acc["fix_rounds"].append(r["rounds"]); acc["fix_nnz"].append(r["nnz_updates"])
        ref.close()
    for k, v in acc.items():
        out[f"{name}.{k}"] = np.array(v, np.int64 if k.endswith(("verdict", "rounds", "nnz")) else np.float64)


def tls4_cases():
    """BASELINE config 1: test_instances/tls4.nl of the reference, read by minotaur_b200/nl_reader.py (the .nl file
    itself is not copied: the fixture holds the flattened instance), node boxes by branching on its integer
    variables; once as read, once with an incumbent of value 10 (the optimum is 8.3)."""
    from minotaur_b200.nl_reader import read_nl
    out = {}
    P = read_nl("/root/reference/test_instances/tls4.nl")
    lin, tp = P.lin, P.tapes
    assert (P.n_var, P.n_con, tp.n_cons, lin.m) == (105, 64, 4, 60)
    lbs, ubs = branch_boxes(lin.lb, lin.ub, lin.var_type, 12, seed=4, max_depth=6)
    lbs[0], ubs[0] = lin.lb, lin.ub
    obj = (lin.cut_col, lin.cut_val)
    lin.cut_col = lin.cut_val = None
    _emit_minlp(out, "tls4", lin, tp, lbs, ubs)
    lin.cut_col, lin.cut_val = obj
    lin.cut_rhs = 10.0 - lin.obj_const
    _emit_minlp(out, "tls4_inc", lin, tp, lbs, ubs)
    return out


def mps_cases():
    """tests/golden/mps/*.mps read by the reference's own Reader::readMps (Reader.cpp:42-473)."""
    from oracle.pyoracle import ref_read_mps
    out = {}
    d = os.path.join(HERE, "mps")
    for f in sorted(os.listdir(d)):
        if f.endswith(".mps"):
            for k, v in ref_read_mps(os.path.join(d, f)).items():
                out[f"{f[:-4]}.{k}"] = v
    return out


if __name__ == "__main__":
    np.savez_compressed(os.path.join(HERE, "mps_cases.npz"), **mps_cases())
    t4 = tls4_cases()
    np.savez_compressed(os.path.join(HERE, "tls4_cases.npz"), **t4)
    print("tls4_cases.npz", os.path.getsize(os.path.join(HERE, "tls4_cases.npz")), "bytes")
    lin = linear_cases()
    np.savez_compressed(os.path.join(HERE, "linear_cases.npz"), **lin)
    nl = nl_cases()
    np.savez_compressed(os.path.join(HERE, "nl_cases.npz"), **nl)
    for f in ("linear_cases.npz", "nl_cases.npz"):
        print(f, os.path.getsize(os.path.join(HERE, f)), "bytes")
