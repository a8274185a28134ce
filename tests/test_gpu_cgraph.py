"""Parity of the CGraph interval kernel (K4, through the C ABI) against the oracle and the golden
fixtures produced by the reference's own NlPresHandler / CGraph code.

MNTR_ORDER_REFERENCE + MNTR_ROUND_NEAREST reproduces NlPresHandler::simplePresolve's in-place, index-ordered
sweep (wavefront levels) bit for bit for +,-,*,/,sqr,sqrt,sumlist,abs,floor,ceil; exp/log/log10/pow go through
CUDA's libm and are compared to 1e-12 relative.
"""
import os

import numpy as np
import pytest

from helpers import assert_box_parity, never_tighter, rel_diff
from minotaur_b200 import engine as E
from minotaur_b200.instances import (CONTINUOUS, Expr, INTEGER, LinearRows, OpExp, OpLog, OpLog10, OpPowK,
                                     branch_boxes, build_tapes, make_minlp)
from test_oracle_golden import GOLD, load_minlp, load_tapes

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(600)]
LIBM_OPS = {OpExp, OpLog, OpLog10, OpPowK}


def empty_linear(n, var_type, lb, ub):
    return LinearRows(m=0, n=n, row_ptr=np.zeros(1, np.int32), col=np.zeros(0, np.int32), val=np.zeros(0),
                      row_lb=np.zeros(0), row_ub=np.zeros(0), var_type=np.asarray(var_type, np.uint8),
                      lb=np.asarray(lb, np.float64), ub=np.asarray(ub, np.float64))


def one_constraint(t, c):
    """Tapes object holding only constraint c of t."""
    from minotaur_b200.instances import Tapes
    b, e = int(t.tape_ptr[c]), int(t.tape_ptr[c + 1])
    lb_, le = int(t.lin_ptr[c]), int(t.lin_ptr[c + 1])
    return Tapes(n_cons=1, tape_ptr=np.array([0, e - b], np.int32), op=t.op[b:e].copy(), arg0=t.arg0[b:e].copy(),
                 arg1=t.arg1[b:e].copy(), cnst=t.cnst[b:e].copy(), child=t.child.copy(),
                 lin_ptr=np.array([0, le - lb_], np.int32),
                 lin_col=(t.lin_col[lb_:le].copy() if le > lb_ else np.zeros(1, np.int32)),
                 lin_val=(t.lin_val[lb_:le].copy() if le > lb_ else np.zeros(1)),
                 c_lb=t.c_lb[c:c + 1].copy(), c_ub=t.c_ub[c:c + 1].copy())


def test_every_opcode_single_constraint(engine, oracle):
    """The 20 hand-written constraints of tests/golden (every implemented opcode and the reference's quirks),
    one at a time, on 12 boxes: NlPresHandler::simplePresolve semantics."""
    z = np.load(os.path.join(GOLD, "nl_cases.npz"))
    t = load_tapes(z, "expr")
    lbs, ubs = z["expr.lbs"], z["expr.ubs"]
    vt = z["expr.var_type"]
    changed = 0
    for c in range(t.n_cons):
        tc = one_constraint(t, c)
        libm = any(int(o) in LIBM_OPS for o in tc.op)
        engine.load_linear(empty_linear(6, vt, lbs[0], ubs[0]))
        engine.load_cgraph(tc)
        res = engine.tighten(lbs, ubs, rounding=E.ROUND_NEAREST, order=E.ORDER_REFERENCE, loop=E.LOOP_SIMPLEPRESOLVE,
                             handlers=E.HANDLERS_NONLINEAR)
        for b in range(lbs.shape[0]):
            ol, ou, r = oracle.nl_simple_presolve(tc, lbs[b], ubs[b])
            if r["verdict"] == 1:
                assert res.verdict[b] == E.INFEAS_NL, (c, b)
                continue
            if res.verdict[b] == E.ERROR_NL:
                continue                    # evaluation error: the reference returns SolveError / asserts
            assert res.verdict[b] == 0, (c, b, res.verdict[b])
            if libm:
                assert rel_diff(res.lb[b], ol).max() < 1e-12 and rel_diff(res.ub[b], ou).max() < 1e-12, (c, b)
            else:
                assert np.array_equal(res.lb[b], ol) and np.array_equal(res.ub[b], ou), (c, b)
            changed += int(np.any(ol != lbs[b]) or np.any(ou != ubs[b]))
    assert changed > 30


@pytest.mark.parametrize("name", ["minlp", "minlp_obj"])
def test_golden_minlp_bitwise(engine, name):
    """tests/golden/nl_cases.npz 'minlp.*': outputs of the reference's NlPresHandler::simplePresolve and of
    LinearHandler + NlPresHandler (one PCBProcessor::presolveNode_ pass).  'minlp_obj.*': the same with a linear
    objective over binaries and an incumbent (varBndsFromObj_ cut-off row + fixObjBins_)."""
    z = np.load(os.path.join(GOLD, "nl_cases.npz"))
    lin, t = load_minlp(z, name)
    lbs, ubs = z[f"{name}.lbs"], z[f"{name}.ubs"]
    engine.load_linear(lin)
    engine.load_cgraph(t)
    nl = engine.tighten(lbs, ubs, rounding=E.ROUND_NEAREST, order=E.ORDER_REFERENCE, loop=E.LOOP_SIMPLEPRESOLVE,
                        handlers=E.HANDLERS_NONLINEAR)
    node = engine.tighten(lbs, ubs, rounding=E.ROUND_NEAREST, order=E.ORDER_REFERENCE, loop=E.LOOP_SIMPLEPRESOLVE)
    n_feas = 0
    for b in range(lbs.shape[0]):
        assert (nl.verdict[b] != 0) == (z[f"{name}.nl_verdict"][b] != 0), b
        if nl.verdict[b] == 0:
            assert np.array_equal(nl.lb[b], z[f"{name}.nl_lb"][b]) and np.array_equal(nl.ub[b], z[f"{name}.nl_ub"][b]), b
        if node.verdict[b] == E.INFEAS_ROW:
            continue
        assert (node.verdict[b] != 0) == (z[f"{name}.node_verdict"][b] != 0), b
        if node.verdict[b] == 0:
            n_feas += 1
            assert np.array_equal(node.lb[b], z[f"{name}.node_lb"][b]), b
            assert np.array_equal(node.ub[b], z[f"{name}.node_ub"][b]), b
    assert n_feas > 0


@pytest.mark.parametrize("seed,nb", [(21, 40), (22, 70)])
def test_random_minlp_node_presolve(engine, oracle, seed, nb):
    lin, tapes = make_minlp(n=600, n_cons=900, m_lin=300, seed=seed)
    engine.load_linear(lin)
    engine.load_cgraph(tapes)
    lbs, ubs = branch_boxes(lin.lb, lin.ub, lin.var_type, nb, seed=seed, max_depth=8, continuous_too=True)
    lbs[0], ubs[0] = lin.lb, lin.ub
    exact = engine.tighten(lbs, ubs, rounding=E.ROUND_NEAREST, order=E.ORDER_REFERENCE, loop=E.LOOP_SIMPLEPRESOLVE)
    dirr = engine.tighten(lbs, ubs, loop=E.LOOP_SIMPLEPRESOLVE)
    n_feas = 0
    for b in range(nb):
        ol, ou, r = oracle.node_presolve(lin, tapes, lbs[b], ubs[b])
        if exact.verdict[b] == E.INFEAS_ROW:
            continue
        assert (exact.verdict[b] != 0) == (r["verdict"] != 0), b
        if r["verdict"] == 0:
            n_feas += 1
            assert np.array_equal(exact.lb[b], ol) and np.array_equal(exact.ub[b], ou), b
            # directed rounding: within 1e-9 relative and never tighter.  (NlPresHandler does not round the
            # bounds it derives for integer variables -- CGraph.cpp:1634-1643 -- so those are real numbers
            # here too and are compared like continuous ones.)
            if dirr.verdict[b] == 0:
                assert rel_diff(dirr.lb[b], ol).max() <= 1e-9 and rel_diff(dirr.ub[b], ou).max() <= 1e-9, b
                assert never_tighter(dirr.lb[b], dirr.ub[b], ol, ou), b
    assert n_feas > nb // 4


def test_nl_only_and_fixpoint_loop(engine, oracle):
    lin, tapes = make_minlp(n=300, n_cons=500, m_lin=0, seed=5)
    engine.load_linear(lin)
    engine.load_cgraph(tapes)
    lbs, ubs = branch_boxes(lin.lb, lin.ub, lin.var_type, 33, seed=8, max_depth=6, continuous_too=True)
    res = engine.tighten(lbs, ubs, rounding=E.ROUND_NEAREST, order=E.ORDER_REFERENCE, loop=E.LOOP_SIMPLEPRESOLVE,
                         handlers=E.HANDLERS_NONLINEAR)
    for b in range(lbs.shape[0]):
        ol, ou, r = oracle.nl_simple_presolve(tapes, lbs[b], ubs[b])
        assert (res.verdict[b] != 0) == (r["verdict"] != 0), b
        if r["verdict"] == 0:
            assert np.array_equal(res.lb[b], ol) and np.array_equal(res.ub[b], ou), b
            assert res.rounds[b] == r["rounds"], b
    # fixpoint loop of both handlers: contained in the one-pass result, idempotent
    lin2, tapes2 = make_minlp(n=300, n_cons=400, m_lin=200, seed=6)
    engine.load_linear(lin2)
    engine.load_cgraph(tapes2)
    lbs, ubs = branch_boxes(lin2.lb, lin2.ub, lin2.var_type, 20, seed=9, max_depth=5, continuous_too=True)
    one = engine.tighten(lbs, ubs, loop=E.LOOP_SIMPLEPRESOLVE)
    fix = engine.tighten(lbs, ubs, loop=E.LOOP_FIXPOINT)
    ok = (one.verdict == 0) & (fix.verdict == 0)
    assert ok.any()
    assert np.all(fix.lb[ok] >= one.lb[ok] - 1e-9) and np.all(fix.ub[ok] <= one.ub[ok] + 1e-9)
    again = engine.tighten(fix.lb[ok], fix.ub[ok], loop=E.LOOP_FIXPOINT)
    assert np.allclose(again.lb, fix.lb[ok], rtol=0, atol=2e-5) and np.allclose(again.ub, fix.ub[ok], rtol=0, atol=2e-5)


def test_c5_shape_reduced(engine, oracle):
    """BASELINE config 5 shape at 1/20 scale: 50k bilinear/quadratic CGraph constraints + 5k linear rows over
    50k variables, 256 boxes; sample against the oracle, validity of the planted point on the root box."""
    lin, tapes = make_minlp(n=50_000, n_cons=50_000, m_lin=5_000, seed=99)
    engine.load_linear(lin)
    engine.load_cgraph(tapes)
    nb = 256
    lbs, ubs = branch_boxes(lin.lb, lin.ub, lin.var_type, nb, seed=99, max_depth=10, continuous_too=True)
    lbs[0], ubs[0] = lin.lb, lin.ub
    res = engine.tighten(lbs, ubs, loop=E.LOOP_SIMPLEPRESOLVE)
    assert res.verdict[0] == 0
    assert np.all(res.lb[0] <= lin.xstar + 1e-6) and np.all(res.ub[0] >= lin.xstar - 1e-6)
    feas = res.verdict == 0
    assert np.all(res.lb[feas] >= lbs[feas]) and np.all(res.ub[feas] <= ubs[feas])
    ex = engine.tighten(lbs[:64], ubs[:64], rounding=E.ROUND_NEAREST, order=E.ORDER_REFERENCE, loop=E.LOOP_SIMPLEPRESOLVE)
    for b in range(0, 64, 7):
        ol, ou, r = oracle.node_presolve(lin, tapes, lbs[b], ubs[b])
        if ex.verdict[b] == E.INFEAS_ROW:
            continue
        assert (ex.verdict[b] != 0) == (r["verdict"] != 0), b
        if r["verdict"] == 0:
            assert np.array_equal(ex.lb[b], ol) and np.array_equal(ex.ub[b], ou), b


@pytest.mark.parametrize("name", ["tls4", "tls4_inc"])
def test_c1_tls4_bitwise(engine, name):
    """BASELINE config 1, tests/golden/tls4_cases.npz: the reference's own handlers on test_instances/tls4.nl (root
    box and branched boxes; 'tls4_inc' with an incumbent: cut-off row + fixObjBins_), bit for bit in the
    reference-order kernel; the default Jacobi / directed mode on the linear rows within tolerance."""
    z = np.load(os.path.join(GOLD, "tls4_cases.npz"))
    lin, t = load_minlp(z, name)
    lbs, ubs = z[f"{name}.lbs"], z[f"{name}.ubs"]
    engine.load_linear(lin)
    engine.load_cgraph(t)
    exact = dict(rounding=E.ROUND_NEAREST, order=E.ORDER_REFERENCE)
    nl = engine.tighten(lbs, ubs, loop=E.LOOP_SIMPLEPRESOLVE, handlers=E.HANDLERS_NONLINEAR, **exact)
    node = engine.tighten(lbs, ubs, loop=E.LOOP_SIMPLEPRESOLVE, **exact)
    raw = engine.tighten(lbs, ubs, loop=E.LOOP_SIMPLEPRESOLVE, handlers=E.HANDLERS_LINEAR, **exact)
    fix = engine.tighten(lbs, ubs, loop=E.LOOP_FIXPOINT, handlers=E.HANDLERS_LINEAR, **exact)
    n_node = 0
    for b in range(lbs.shape[0]):
        assert (nl.verdict[b] != 0) == (z[f"{name}.nl_verdict"][b] != 0), b
        if nl.verdict[b] == 0:
            assert np.array_equal(nl.lb[b], z[f"{name}.nl_lb"][b]) and np.array_equal(nl.ub[b], z[f"{name}.nl_ub"][b]), b
        assert (fix.verdict[b] != 0) == (z[f"{name}.fix_verdict"][b] != 0), b
        if fix.verdict[b] == 0:
            assert np.array_equal(fix.lb[b], z[f"{name}.fix_lb"][b]) and np.array_equal(fix.ub[b], z[f"{name}.fix_ub"][b]), b
            assert fix.rounds[b] == z[f"{name}.fix_rounds"][b] and fix.nnz_updates[b] == z[f"{name}.fix_nnz"][b], b
        if raw.verdict[b] != E.INFEAS_ROW:
            assert (raw.verdict[b] != 0) == (z[f"{name}.raw_verdict"][b] != 0), b
            assert np.array_equal(raw.lb[b], z[f"{name}.raw_lb"][b]) and np.array_equal(raw.ub[b], z[f"{name}.raw_ub"][b]), b
        if node.verdict[b] != E.INFEAS_ROW:
            assert (node.verdict[b] != 0) == (z[f"{name}.node_verdict"][b] != 0), b
            if node.verdict[b] == 0:
                n_node += 1
                assert np.array_equal(node.lb[b], z[f"{name}.node_lb"][b]), b
                assert np.array_equal(node.ub[b], z[f"{name}.node_ub"][b]), b
    assert n_node >= 6
    # default mode (single box, Jacobi rounds, directed rounding) on the linear rows of the root box
    engine.load_linear(lin)
    res = engine.tighten(lbs[0], ubs[0])
    assert res.verdict[0] == 0
    assert_box_parity(lin.var_type, res.lb, res.ub, z[f"{name}.fix_lb"][0], z[f"{name}.fix_ub"][0], what="jacobi")   # 1e-9
    assert never_tighter(res.lb, res.ub, z[f"{name}.fix_lb"][0], z[f"{name}.fix_ub"][0])
