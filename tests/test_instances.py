"""CPU tests of the bench-size generators and of the oracle's batch driver (test infrastructure for the in-run parity
checks of bench.py)."""
import numpy as np

from minotaur_b200.instances import (BINARY, INTEGER, branch_deltas, build_tapes, deltas_box, make_knapsack_setcover,
                                     make_minlp_large, make_sparse_milp, minlp_tapes_from_draws, slice_deltas)


def test_minlp_tapes_direct_layout_equals_build_tapes():
    """make_minlp_large writes the tapes directly; they must be the tapes flatten_expr / build_tapes (the node order of
    CGraph::finalize, pinned against the reference in test_tape_order_matches_cgraph_finalize) produce."""
    rng = np.random.default_rng(5)
    n = 60
    ijk = np.array([rng.choice(n, 3, replace=False) for _ in range(64)])
    xs = rng.random(n) * 4 - 2
    t, cons = minlp_tapes_from_draws(ijk, xs, np.random.default_rng(1), return_cons=True)
    t2 = build_tapes(cons)
    for f in ("tape_ptr", "op", "arg0", "arg1", "cnst", "child", "lin_ptr", "lin_col", "lin_val", "c_lb", "c_ub"):
        a, b = getattr(t, f), getattr(t2, f)
        assert a.shape == b.shape and np.array_equal(a, b), f


def test_minlp_large_root_is_feasible_for_the_reference_rules(oracle):
    lin, tapes = make_minlp_large(4000, 4000, 400)
    assert np.all(lin.lb < 0) and np.all(lin.ub > 0)                     # no zero end points at the root
    assert np.all(lin.lb <= lin.xstar) and np.all(lin.xstar <= lin.ub)
    _, _, r = oracle.node_presolve(lin, tapes, lin.lb, lin.ub)
    assert r["verdict"] == 0


def test_branch_deltas_are_branching_perturbations():
    inst = make_knapsack_setcover(3000, 3000, 10, seed=3)
    ptr, var, up, val = d = branch_deltas(inst.lb, inst.ub, inst.var_type, 200, seed=4, max_depth=20)
    assert ptr[0] == 0 and ptr[-1] == len(var) and np.all(np.diff(ptr) >= 0) and np.all(np.diff(ptr) <= 20)
    isint = (inst.var_type == INTEGER) | (inst.var_type == BINARY)
    assert np.all(isint[var]) and np.all(val == np.round(val))
    for b in (0, 17, 199):
        lb, ub = deltas_box(inst.lb, inst.ub, d, b)
        assert np.all(lb >= inst.lb) and np.all(ub <= inst.ub) and np.all(lb <= ub)
        q = slice(int(ptr[b]), int(ptr[b + 1]))
        assert len(set(var[q].tolist())) == ptr[b + 1] - ptr[b]              # one perturbation per variable
    sub = slice_deltas(d, 10, 20)
    assert len(sub[0]) == 11 and sub[0][-1] == ptr[20] - ptr[10]


def test_oracle_batch_driver_equals_box_by_box(oracle):
    inst = make_sparse_milp(600, 500, 6, seed=9)
    d = branch_deltas(inst.lb, inst.ub, inst.var_type, 40, seed=2, max_depth=8)
    one = oracle.batch_deltas(inst, None, 0, inst.lb, inst.ub, d, n_threads=1, mod_cap=2048)
    par = oracle.batch_deltas(inst, None, 0, inst.lb, inst.ub, d, n_threads=4, mod_cap=2048)
    for k in ("verdict", "rounds", "nnz", "mod_cnt", "mod_var", "mod_up", "mod_val"):
        assert np.array_equal(one[k], par[k]), k
    raw = oracle.batch_deltas(inst, None, 1, inst.lb, inst.ub, d, n_threads=2, mod_cap=2048)
    for b in range(40):
        lb, ub = deltas_box(inst.lb, inst.ub, d, b)
        l, u, r = oracle.lin_fixpoint_inplace(inst, lb, ub)
        assert r["verdict"] == one["verdict"][b] and r["nnz_updates"] == one["nnz"][b] and r["rounds"] == one["rounds"][b]
        if r["verdict"] == 0:
            c = one["mod_cnt"][b]
            L, U = lb.copy(), ub.copy()
            for q in range(c):
                (U if one["mod_up"][b, q] else L)[one["mod_var"][b, q]] = one["mod_val"][b, q]
            assert np.array_equal(L, l) and np.array_equal(U, u)
        l2, u2, r2 = oracle.lin_simple_presolve(inst, lb, ub)
        assert r2["verdict"] == raw["verdict"][b] and r2["nnz_updates"] == raw["nnz"][b]
