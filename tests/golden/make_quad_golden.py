"""Writes tests/golden/quad_cases.npz: QuadraticFunction constraints checked by the REFERENCE's own
NlPresHandler::chkRed_ / QuadraticFunction::computeBounds (oracle/_ref, built from /root/reference) on random boxes.
Run where /root/reference exists:  python tests/golden/make_quad_golden.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from minotaur_b200.instances import branch_boxes, make_minlp, make_quad_cons  # noqa: E402
from oracle import pyoracle  # noqa: E402

out = {}
names = []
for name, n, n_cons, n_quad, seed in (("q_only", 80, 0, 40, 1), ("q_mixed", 120, 30, 50, 2), ("q_wide", 60, 10, 30, 3)):
    lin, tapes = make_minlp(n=n, n_cons=max(n_cons, 1), m_lin=20, seed=seed)
    if n_cons == 0:
        tapes = None
    quad = make_quad_cons(lin.n, n_quad, lin.xstar, seed=seed, terms=6 if name == "q_wide" else 4)
    ref = pyoracle.Reference(lin, tapes, quad)
    lbs, ubs = branch_boxes(lin.lb, lin.ub, lin.var_type, 48, seed=seed + 10, max_depth=14, continuous_too=True)
    lbs[0], ubs[0] = lin.lb, lin.ub
    if name == "q_wide":        # some unbounded variables: infinities and 0 * inf in the corner products
        lbs[1::3, ::7] = -np.inf
        ubs[2::3, ::5] = np.inf
    verdict = np.array([ref.nl_chk_red(lbs[b], ubs[b]) for b in range(lbs.shape[0])], np.int32)
    qb = np.array([[ref.quad_compute_bounds(q, lbs[b], ubs[b]) for q in range(quad.n_quad)] for b in range(lbs.shape[0])])
    ref.close()
    names.append(name)
    for k in ("row_ptr", "col", "val", "row_lb", "row_ub", "var_type", "lb", "ub"):
        out[f"{name}.{k}"] = getattr(lin, k)
    out[f"{name}.has_tapes"] = np.array([tapes is not None])
    if tapes is not None:
        for k in ("tape_ptr", "op", "arg0", "arg1", "cnst", "child", "lin_ptr", "lin_col", "lin_val", "c_lb", "c_ub"):
            out[f"{name}.t.{k}"] = getattr(tapes, k)
    for k in ("q_ptr", "v1", "v2", "coef", "lin_ptr", "lin_col", "lin_val", "q_lb", "q_ub"):
        out[f"{name}.q.{k}"] = getattr(quad, k)
    out[f"{name}.lbs"], out[f"{name}.ubs"] = lbs, ubs
    out[f"{name}.chk_verdict"] = verdict
    out[f"{name}.quad_bounds"] = qb
    print(name, "infeasible boxes:", int(verdict.sum()), "of", len(verdict))
out["names"] = np.array(names)
np.savez_compressed(os.path.join(HERE, "quad_cases.npz"), **out)
