"""Live pin: oracle/fbbt_oracle.c against the reference's own objects (oracle/_ref), on seeded random
instances.  Skipped where the reference library was not built (it needs /root/reference)."""
import numpy as np
import pytest

from minotaur_b200.instances import attach_binary_objective, attach_cutoff, branch_boxes, make_knapsack_setcover, make_minlp, make_sparse_milp
from helpers import assert_box_parity


@pytest.fixture(scope="module")
def Reference(have_ref):
    if not have_ref:
        pytest.skip("oracle/_ref/libminotaur_ref.so not built")
    from oracle.pyoracle import Reference
    return Reference


@pytest.mark.parametrize("seed,real,inf", [(0, False, (0, 0, 0)), (1, True, (0, 0, 0)), (2, True, (0.1, 0.05, 0.02)),
                                           (3, False, (0.2, 0.1, 0.1))])
def test_linear_inplace_bitwise(oracle, Reference, seed, real, inf):
    inst = make_sparse_milp(250, 220, 6, seed=seed, real_data=real, inf_frac=inf)
    ref = Reference(inst)
    lbs, ubs = branch_boxes(inst.lb, inst.ub, inst.var_type, 25, seed=seed, max_depth=10)
    n_inf = 0
    for b in range(lbs.shape[0]):
        rl, ru, rr = ref.lin_simple_presolve(lbs[b], ubs[b])
        ol, ou, orr = oracle.lin_simple_presolve(inst, lbs[b], ubs[b])
        assert rr["verdict"] == orr["verdict"] and rr["n_mods"] == orr["n_mods"]
        assert np.array_equal(rl, ol) and np.array_equal(ru, ou)
        rl, ru, rr = ref.lin_fixpoint(lbs[b], ubs[b], counted=True)
        ol, ou, orr = oracle.lin_fixpoint_inplace(inst, lbs[b], ubs[b])
        assert rr["verdict"] == orr["verdict"] and rr["rounds"] == orr["rounds"]
        assert rr["nnz_updates"] == orr["nnz_updates"]
        n_inf += rr["verdict"]
        if rr["verdict"] == 0:
            assert np.array_equal(rl, ol) and np.array_equal(ru, ou)
            jl, ju, jr = oracle.lin_fixpoint_jacobi(inst, lbs[b], ubs[b])
            assert jr["verdict"] == 0
            assert_box_parity(inst.var_type, jl, ju, rl, ru, what=f"jacobi box {b}")       # 1e-9
        else:
            assert oracle.lin_fixpoint_jacobi(inst, lbs[b], ubs[b])[2]["verdict"] != 0
    ref.close()


@pytest.mark.parametrize("seed,real,inf,k,slack", [(10, False, (0, 0, 0), 40, 20.0), (11, True, (0, 0, 0), 90, 30.0),
                                                   (12, True, (0.1, 0.05, 0.02), 60, 20.0), (13, False, (0, 0, 0), 220, 30.0)])
def test_linear_cutoff_row_bitwise(oracle, Reference, seed, real, inf, k, slack):
    """Objective cut-off row (LinearHandler::varBndsFromObj_, :544-597): the reference meets it as a linear objective
    plus an incumbent in the solution pool; raw simplePresolve and the status-honouring fixpoint, bit for bit."""
    inst = make_sparse_milp(250, 220, 6, seed=seed, real_data=real, inf_frac=inf)
    tl, tu, _ = oracle.lin_fixpoint_inplace(inst, inst.lb, inst.ub)
    inst = attach_cutoff(inst, k, seed, slack, box=(tl, tu))
    ref = Reference(inst)
    lbs, ubs = branch_boxes(inst.lb, inst.ub, inst.var_type, 16, seed=seed, max_depth=8)
    lbs[0], ubs[0] = inst.lb, inst.ub
    n_feas = moved = 0
    for b in range(lbs.shape[0]):
        rl, ru, rr = ref.lin_simple_presolve(lbs[b], ubs[b])
        ol, ou, orr = oracle.lin_simple_presolve(inst, lbs[b], ubs[b])
        assert rr["verdict"] == orr["verdict"] and rr["n_mods"] == orr["n_mods"], b
        assert np.array_equal(rl, ol) and np.array_equal(ru, ou), b
        rl, ru, rr = ref.lin_fixpoint(lbs[b], ubs[b], counted=True)
        ul, uu, ur = ref.lin_fixpoint(lbs[b], ubs[b], counted=False)         # the reference's own varBndsFromObj_
        ol, ou, orr = oracle.lin_fixpoint_inplace(inst, lbs[b], ubs[b])
        assert rr["verdict"] == ur["verdict"] == orr["verdict"] and rr["rounds"] == ur["rounds"] == orr["rounds"], b
        assert rr["nnz_updates"] == orr["nnz_updates"], b
        if rr["verdict"] == 0:
            n_feas += 1
            assert np.array_equal(rl, ol) and np.array_equal(ru, ou) and np.array_equal(ul, ol) and np.array_equal(uu, ou), b
            # what the row did: the same box without an incumbent
            ref.set_incumbent(None)
            nl, nu, nr = ref.lin_fixpoint(lbs[b], ubs[b])
            ref.set_incumbent(inst.cut_rhs)
            moved += int(np.sum(rl > nl) + np.sum(ru < nu))
    ref.close()
    assert n_feas > 0 and moved > 0, "the cut-off row never moved a bound: vacuous"


def test_knapsack_setcover_bitwise(oracle, Reference):
    inst = make_knapsack_setcover(300, 200, 6, seed=4)
    ref = Reference(inst)
    lbs, ubs = branch_boxes(inst.lb, inst.ub, inst.var_type, 30, seed=9, max_depth=12)
    for b in range(lbs.shape[0]):
        rl, ru, rr = ref.lin_fixpoint(lbs[b], ubs[b], counted=True)
        ol, ou, orr = oracle.lin_fixpoint_inplace(inst, lbs[b], ubs[b])
        assert rr["verdict"] == orr["verdict"] and rr["nnz_updates"] == orr["nnz_updates"]
        if rr["verdict"] == 0:
            assert np.array_equal(rl, ol) and np.array_equal(ru, ou)
    ref.close()


def test_minlp_node_presolve_bitwise(oracle, Reference):
    lin, tapes = make_minlp(n=60, n_cons=90, m_lin=30, seed=21)
    lbs, ubs = branch_boxes(lin.lb, lin.ub, lin.var_type, 10, seed=5, max_depth=6, continuous_too=True)
    for b in range(lbs.shape[0]):
        ref = Reference(lin, tapes)                 # fresh graph: constant nodes keep state in the reference
        rl, ru, rr = ref.node_presolve(lbs[b], ubs[b])
        ol, ou, orr = oracle.node_presolve(lin, tapes, lbs[b], ubs[b])
        assert (rr["verdict"] != 0) == (orr["verdict"] != 0)
        if rr["verdict"] == 0:
            assert np.array_equal(rl, ol) and np.array_equal(ru, ou)
        ref.close()


@pytest.mark.parametrize("seed,slack,const", [(31, 4.0, 0.0), (35, 3.0, 2.25), (36, 9.0, -4.5)])
def test_fix_obj_bins_bitwise(oracle, Reference, seed, slack, const):
    """NlPresHandler::fixObjBins_ (NlPresHandler.cpp:1062-1121): with an incumbent, binaries of the linear
    objective whose coefficient alone exceeds the slack are fixed.  The rule compares against the RAW incumbent
    value (the objective constant is not subtracted, :1030), hence the cases with a constant."""
    lin, tapes = make_minlp(n=300, n_cons=200, m_lin=60, seed=seed)
    lin = attach_binary_objective(lin, n_bin=14, n_other=2, seed=seed, slack=slack, const=const, coef_hi=40)
    lbs, ubs = branch_boxes(lin.lb, lin.ub, lin.var_type, 8, seed=seed, max_depth=5, continuous_too=True)
    lbs[0], ubs[0] = lin.lb, lin.ub
    fixed = 0
    for b in range(lbs.shape[0]):
        ref = Reference(lin, tapes)                 # fresh graph: constant nodes keep state in the reference
        rl, ru, rr = ref.nl_simple_presolve(lbs[b], ubs[b])
        ol, ou, orr = oracle.nl_simple_presolve(tapes, lbs[b], ubs[b], obj=lin)
        assert (rr["verdict"] == 1) == (orr["verdict"] == 1), b
        if rr["verdict"] == 0:
            assert np.array_equal(rl, ol) and np.array_equal(ru, ou), b
            assert rr["n_mods"] == orr["n_mods"], b
            nl, nu, _ = oracle.nl_simple_presolve(tapes, lbs[b], ubs[b])       # without the incumbent
            fixed += int(np.sum(ol > nl) + np.sum(ou < nu))
        ref.close()
        ref = Reference(lin, tapes)
        rl, ru, rr = ref.node_presolve(lbs[b], ubs[b])
        ol, ou, orr = oracle.node_presolve(lin, tapes, lbs[b], ubs[b])
        assert (rr["verdict"] != 0) == (orr["verdict"] != 0), b
        if rr["verdict"] == 0:
            assert np.array_equal(rl, ol) and np.array_equal(ru, ou), b
        ref.close()
    assert fixed > 0, "fixObjBins_ never fixed a binary: vacuous"
