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
