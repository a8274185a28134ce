"""Shared comparison helpers of the parity tests."""
import numpy as np

from minotaur_b200.instances import BINARY, INTEGER

REL_TOL = 1e-9      # north star: continuous bounds agree within 1e-9 relative
INT_TOL = 1e-6      # LinearHandler::intTol_


def is_int_var(var_type):
    return (var_type == INTEGER) | (var_type == BINARY)


def canon_int(x):
    """Integer-variable bounds are compared after the canonicalisation of SURVEY.md section 7 (hard
    part 2): the reference leaves a bound within 1e-6 of an integer as is, so a value such as
    2.9999999999999996 and 3.0 denote the same integer bound."""
    r = np.round(x)
    with np.errstate(invalid="ignore"):
        return np.where(np.isfinite(x) & (np.abs(x - r) <= INT_TOL), r, x)


def rel_diff(a, b):
    with np.errstate(invalid="ignore"):
        d = np.abs(a - b) / np.maximum(1.0, np.maximum(np.abs(a), np.abs(b)))
    d = np.where(a == b, 0.0, d)           # equal infinities
    return np.nan_to_num(d, nan=np.inf)


def assert_box_parity(var_type, got_lb, got_ub, ref_lb, ref_ub, rel_tol=REL_TOL, exact=False, what=""):
    """got = CUDA path, ref = oracle/reference.  Integer-variable bounds bit-exact (after canon_int unless
    exact), continuous within rel_tol and never tighter than the reference beyond rel_tol."""
    isint = is_int_var(var_type)
    if exact:
        assert np.array_equal(got_lb, ref_lb), f"{what}: lb differs bitwise at {np.nonzero(got_lb != ref_lb)[0][:10]}"
        assert np.array_equal(got_ub, ref_ub), f"{what}: ub differs bitwise at {np.nonzero(got_ub != ref_ub)[0][:10]}"
        return
    gi_l, ri_l = canon_int(got_lb[isint]), canon_int(ref_lb[isint])
    gi_u, ri_u = canon_int(got_ub[isint]), canon_int(ref_ub[isint])
    assert np.array_equal(gi_l, ri_l), f"{what}: integer lb differ at {np.nonzero(gi_l != ri_l)[0][:10]}"
    assert np.array_equal(gi_u, ri_u), f"{what}: integer ub differ at {np.nonzero(gi_u != ri_u)[0][:10]}"
    c = ~isint
    dl, du = rel_diff(got_lb[c], ref_lb[c]), rel_diff(got_ub[c], ref_ub[c])
    assert dl.max(initial=0.0) <= rel_tol, f"{what}: continuous lb rel diff {dl.max()}"
    assert du.max(initial=0.0) <= rel_tol, f"{what}: continuous ub rel diff {du.max()}"


def never_tighter(got_lb, got_ub, ref_lb, ref_ub, rel_tol=REL_TOL):
    """True when the GPU box contains the reference box up to rel_tol."""
    sl = np.maximum(1.0, np.abs(ref_lb)) * rel_tol
    su = np.maximum(1.0, np.abs(ref_ub)) * rel_tol
    with np.errstate(invalid="ignore"):
        ok_l = (got_lb <= ref_lb + sl) | (got_lb == ref_lb)
        ok_u = (got_ub >= ref_ub - su) | (got_ub == ref_ub)
    return bool(np.all(ok_l) and np.all(ok_u))


def ragged_rows(inst, seed, max_keep=None):
    """A copy of a uniform-row instance (k entries per row) whose rows keep a random number 0..max_keep of their
    entries (in column order), row bounds rebuilt around the planted point: ragged blocks for the row kernels."""
    import copy
    rng = np.random.default_rng(seed)
    m = inst.m
    k = inst.nnz // m
    assert inst.nnz == m * k and inst.xstar is not None
    keep = rng.integers(0, (max_keep or k) + 1, size=m)
    # a few long rows next to many short ones
    keep = np.where(rng.random(m) < 0.85, np.minimum(keep, rng.integers(0, 14, size=m)), keep)
    mask = np.arange(k)[None, :] < keep[:, None]
    col = inst.col.reshape(m, k)[mask].astype(np.int32)
    val = inst.val.reshape(m, k)[mask]
    row_ptr = np.concatenate([[0], np.cumsum(keep)]).astype(np.int32)
    act = np.add.reduceat(np.concatenate([val * inst.xstar[col], [0.0]]), np.minimum(row_ptr[:-1], len(col)))[:m]
    act = np.where(keep > 0, act, 0.0)
    is_eq = inst.row_lb == inst.row_ub
    slack = rng.integers(0, 4, size=m).astype(np.float64)
    out = copy.copy(inst)
    out.row_ptr, out.col, out.val = row_ptr, col, val
    out.row_ub = np.where(is_eq, act, act + slack)
    out.row_lb = np.where(is_eq, act, -np.inf)
    out.name = inst.name + "-ragged"
    out.validate()
    return out
