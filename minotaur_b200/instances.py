"""Flat problem descriptions and the synthetic instance generators of SURVEY.md section 8(d).

The flat layout is what crosses the C ABI (include/mntr_gpu.h): a CSR matrix of the
linear rows with columns in ascending variable id (the iteration order of the
reference's ``LinearFunction`` terms, /root/reference/src/base/Types.h:496), row and
variable bounds, Minotaur ``VariableType`` codes (Types.h:83-89), and one expression
tape per CGraph constraint in the evaluation order ``CGraph::finalize`` produces
(/root/reference/src/base/CGraph.cpp:557-644).

Pure numpy: nothing here touches the GPU or the oracle.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

import numpy as np

# Minotaur::VariableType (Types.h:83-89)
BINARY, INTEGER, IMPLBIN, IMPLINT, CONTINUOUS = 0, 1, 2, 3, 4

# Minotaur::OpCode (OpCode.h:17-53)
(OpAbs, OpAcos, OpAcosh, OpAsin, OpAsinh, OpAtan, OpAtanh, OpCeil, OpCos, OpCosh, OpCPow, OpDiv,
 OpExp, OpFloor, OpInt, OpIntDiv, OpLog, OpLog10, OpMinus, OpMult, OpNone, OpNum, OpPlus, OpPow,
 OpPowK, OpRound, OpSin, OpSinh, OpSqr, OpSqrt, OpSumList, OpTan, OpTanh, OpUMinus, OpVar) = range(35)

INF = float("inf")


@dataclass
class LinearRows:
    """CSR linear rows  row_lb <= A x <= row_ub  plus variable types and the root box."""
    m: int
    n: int
    row_ptr: np.ndarray   # int32 [m+1]
    col: np.ndarray       # int32 [nnz], ascending inside a row
    val: np.ndarray       # float64 [nnz]
    row_lb: np.ndarray    # float64 [m]
    row_ub: np.ndarray    # float64 [m]
    var_type: np.ndarray  # uint8 [n]
    lb: np.ndarray        # float64 [n] root box
    ub: np.ndarray        # float64 [n]
    row_active: Optional[np.ndarray] = None   # uint8 [m] or None
    cut_col: Optional[np.ndarray] = None      # objective cut-off row  c.x <= cut_rhs
    cut_val: Optional[np.ndarray] = None
    cut_rhs: float = INF
    obj_const: float = 0.0                    # objective constant: the incumbent's value is cut_rhs + obj_const
    name: str = ""
    xstar: Optional[np.ndarray] = None        # planted feasible point of the generators

    @property
    def nnz(self) -> int:
        return int(self.row_ptr[-1])

    def validate(self) -> None:
        assert self.row_ptr.dtype == np.int32 and self.col.dtype == np.int32
        assert self.val.dtype == np.float64 and self.var_type.dtype == np.uint8
        assert self.row_ptr.shape == (self.m + 1,) and self.row_ptr[0] == 0
        assert self.col.shape == (self.nnz,) and self.val.shape == (self.nnz,)
        assert self.lb.shape == (self.n,) and self.ub.shape == (self.n,)
        if self.nnz:
            assert self.col.min() >= 0 and self.col.max() < self.n
        for i in range(min(self.m, 1000)):
            c = self.col[self.row_ptr[i]:self.row_ptr[i + 1]]
            assert np.all(np.diff(c) > 0), "columns must be strictly ascending inside a row"


@dataclass
class Tapes:
    """Expression tapes of the CGraph constraints  c_lb <= f_c(x) + lin_c . x <= c_ub."""
    n_cons: int
    tape_ptr: np.ndarray   # int32 [n_cons+1]
    op: np.ndarray         # uint8 [n_nodes]
    arg0: np.ndarray       # int32 [n_nodes]
    arg1: np.ndarray       # int32 [n_nodes]
    cnst: np.ndarray       # float64 [n_nodes]
    child: np.ndarray      # int32 [n_child]  SumList child lists
    lin_ptr: np.ndarray    # int32 [n_cons+1]
    lin_col: np.ndarray    # int32
    lin_val: np.ndarray    # float64
    c_lb: np.ndarray       # float64 [n_cons]
    c_ub: np.ndarray       # float64 [n_cons]

    @property
    def n_nodes(self) -> int:
        return int(self.tape_ptr[-1])


@dataclass
class QuadCons:
    """Constraints with a QuadraticFunction:  q_lb <= sum_k coef_k x_{v1_k} x_{v2_k} + lin . x <= q_ub.  Terms are kept
    in the order of the reference's VariablePairGroup (by (v1, v2) with v1 <= v2,
    /root/reference/src/base/Types.cpp CompareVariablePair); NlPresHandler::chkRed_ checks them with
    QuadraticFunction::computeBounds (QuadraticFunction.cpp:156-180)."""
    n_quad: int
    q_ptr: np.ndarray      # int32 [n_quad+1]
    v1: np.ndarray         # int32
    v2: np.ndarray         # int32
    coef: np.ndarray       # float64
    lin_ptr: np.ndarray    # int32 [n_quad+1]
    lin_col: np.ndarray    # int32
    lin_val: np.ndarray    # float64
    q_lb: np.ndarray       # float64 [n_quad]
    q_ub: np.ndarray       # float64 [n_quad]


def build_quad(cons) -> QuadCons:
    """cons: list of ([(i, j, coef) ...], [(col, coef) ...], lb, ub).  Terms are merged per variable pair and sorted the
    way QuadraticFunction keeps them (a std::map keyed by the ordered pair; coefficients below 1e-8 are dropped,
    QuadraticFunction.cpp addTerm)."""
    qp, lp = [0], [0]
    v1, v2, cf, lc, lv, lo, hi = [], [], [], [], [], [], []
    for terms, lin, l, u in cons:
        seen = {}
        for i, j, c in terms:
            key = (min(i, j), max(i, j))
            if abs(c) >= 1e-8 and key not in seen:       # std::map::insert keeps the first weight of a pair
                seen[key] = c
        for key in sorted(seen):
            v1.append(key[0]); v2.append(key[1]); cf.append(seen[key])
        qp.append(len(v1))
        lin = sorted(lin)
        lc += [j for j, _ in lin]; lv += [a for _, a in lin]
        lp.append(len(lc))
        lo.append(l); hi.append(u)
    return QuadCons(n_quad=len(cons), q_ptr=np.asarray(qp, np.int32), v1=np.asarray(v1 or [0], np.int32),
                    v2=np.asarray(v2 or [0], np.int32), coef=np.asarray(cf or [0.0], np.float64),
                    lin_ptr=np.asarray(lp, np.int32), lin_col=np.asarray(lc or [0], np.int32),
                    lin_val=np.asarray(lv or [0.0], np.float64), q_lb=np.asarray(lo, np.float64), q_ub=np.asarray(hi, np.float64))


def make_quad_cons(n: int, n_quad: int, xstar: np.ndarray, seed: int = 7, terms: int = 4) -> QuadCons:
    """Random quadratic constraints around a planted point: a few bilinear / square terms with coefficients in +-{1..4},
    one or two linear terms, bounds x*'s value plus a small slack (a third two-sided)."""
    rng = np.random.default_rng([seed, 13])
    cons = []
    for _ in range(n_quad):
        k = int(rng.integers(1, terms + 1))
        tt = []
        val = 0.0
        for _ in range(k):
            i, j = int(rng.integers(0, n)), int(rng.integers(0, n))
            c = float(rng.integers(1, 5)) * (1.0 if rng.random() < 0.6 else -1.0)
            tt.append((i, j, c))
        merged = {}
        for i, j, c in tt:
            merged.setdefault((min(i, j), max(i, j)), c)
        for (i, j), c in merged.items():
            val += c * xstar[i] * xstar[j]
        lin = []
        for _ in range(int(rng.integers(0, 3))):
            j = int(rng.integers(0, n)); a = float(rng.integers(1, 6)) * (1.0 if rng.random() < 0.5 else -1.0)
            if all(j != jj for jj, _ in lin):
                lin.append((j, a)); val += a * xstar[j]
        w = float(rng.integers(0, 4))
        if rng.random() < 0.33:
            cons.append((tt, lin, val - 0.5 * w, val + 0.5 * w))
        elif rng.random() < 0.5:
            cons.append((tt, lin, -INF, val + w))
        else:
            cons.append((tt, lin, val - w, INF))
    return build_quad(cons)


# --------------------------------------------------------------------------------------
# expression -> tape (mirrors the node order of CGraph::finalize)
# --------------------------------------------------------------------------------------

class Expr:
    """Tiny expression DAG used to author CGraph constraints for tests and generators."""
    __slots__ = ("op", "kids", "value", "var")

    def __init__(self, op, kids=(), value=0.0, var=-1):
        self.op, self.kids, self.value, self.var = op, tuple(kids), float(value), int(var)

    @staticmethod
    def v(j):
        return Expr(OpVar, var=j)

    @staticmethod
    def c(x):
        return Expr(OpNum, value=x)

    @staticmethod
    def i(x):
        return Expr(OpInt, value=x)

    def __add__(self, o): return Expr(OpPlus, (self, _e(o)))
    def __sub__(self, o): return Expr(OpMinus, (self, _e(o)))
    def __mul__(self, o): return Expr(OpMult, (self, _e(o)))
    def __truediv__(self, o): return Expr(OpDiv, (self, _e(o)))
    def __neg__(self): return Expr(OpUMinus, (self,))

    def sqr(self): return Expr(OpSqr, (self,))
    def sqrt(self): return Expr(OpSqrt, (self,))
    def exp(self): return Expr(OpExp, (self,))
    def log(self): return Expr(OpLog, (self,))
    def abs(self): return Expr(OpAbs, (self,))
    def powk(self, k): return Expr(OpPowK, (self, Expr.c(k)))

    @staticmethod
    def sumlist(kids): return Expr(OpSumList, tuple(_e(k) for k in kids))

    @staticmethod
    def unary(op, a): return Expr(op, (_e(a),))


def _e(x):
    return x if isinstance(x, Expr) else Expr.c(x)


def flatten_expr(root: Expr) -> Tuple[list, list, list, list, list]:
    """Return (op, arg0, arg1, cnst, child) of one constraint in reference order:
    variable nodes by ascending variable id (``vq_``), then constants, then operator
    nodes in the iterative left-to-right post-order of ``CGraph::finalize`` (``dq_``).
    One node per variable (``CGraph::newNode(VariablePtr)`` de-duplicates,
    CGraph.cpp:1247-1262); shared sub-expressions (same Expr object) are emitted once."""
    var_nodes, consts, ops = {}, [], []
    seen = {}

    def visit(e):
        stack = [(e, 0)]
        while stack:
            node, k = stack.pop()
            if id(node) in seen:
                continue
            if node.op == OpVar:
                var_nodes.setdefault(node.var, node)
                seen[id(node)] = True
                continue
            if node.op in (OpNum, OpInt):
                consts.append(node)
                seen[id(node)] = True
                continue
            if k < len(node.kids):
                stack.append((node, k + 1))
                stack.append((node.kids[k], 0))
            else:
                seen[id(node)] = True
                ops.append(node)

    visit(root)
    order, index = [], {}
    for j in sorted(var_nodes):
        index[("v", j)] = len(order)
        order.append(var_nodes[j])
    for cn in consts:
        index[id(cn)] = len(order)
        order.append(cn)
    for on in ops:
        index[id(on)] = len(order)
        order.append(on)

    def idx(e):
        return index[("v", e.var)] if e.op == OpVar else index[id(e)]

    op, a0, a1, cn, child = [], [], [], [], []
    for e in order:
        op.append(e.op)
        cn.append(e.value if e.op in (OpNum, OpInt) else 0.0)
        if e.op == OpVar:
            a0.append(e.var); a1.append(-1)
        elif e.op in (OpNum, OpInt):
            a0.append(-1); a1.append(-1)
        elif e.op == OpSumList:
            a0.append(len(child))
            child.extend(idx(k) for k in e.kids)
            a1.append(len(child))
        else:
            a0.append(idx(e.kids[0]))
            a1.append(idx(e.kids[1]) if len(e.kids) > 1 else -1)
    if root.op in (OpVar, OpNum, OpInt):
        raise ValueError("a CGraph constraint needs at least one operator node")
    return op, a0, a1, cn, child


def build_tapes(cons: Sequence[Tuple[Expr, Sequence[Tuple[int, float]], float, float]]) -> Tapes:
    """cons: list of (expr, [(col, coef) ...] linear part, c_lb, c_ub)."""
    tp, lp = [0], [0]
    op, a0, a1, cn, child, lc, lv, clb, cub = [], [], [], [], [], [], [], [], []
    for expr, lin, lo, hi in cons:
        o, x0, x1, c, ch = flatten_expr(expr)
        base_child = len(child)
        for k, oo in enumerate(o):
            if oo == OpSumList:
                x0[k] += base_child
                x1[k] += base_child
        op += o; a0 += x0; a1 += x1; cn += c; child += ch
        tp.append(len(op))
        lin = sorted(lin)
        lc += [j for j, _ in lin]; lv += [a for _, a in lin]
        lp.append(len(lc))
        clb.append(lo); cub.append(hi)
    return Tapes(
        n_cons=len(cons), tape_ptr=np.asarray(tp, np.int32), op=np.asarray(op, np.uint8),
        arg0=np.asarray(a0, np.int32), arg1=np.asarray(a1, np.int32), cnst=np.asarray(cn, np.float64),
        child=np.asarray(child if child else [0], np.int32), lin_ptr=np.asarray(lp, np.int32),
        lin_col=np.asarray(lc if lc else [0], np.int32), lin_val=np.asarray(lv if lv else [0.0], np.float64),
        c_lb=np.asarray(clb, np.float64), c_ub=np.asarray(cub, np.float64))


# --------------------------------------------------------------------------------------
# generators
# --------------------------------------------------------------------------------------

def _distinct_sorted_columns(rng: np.random.Generator, m: int, n: int, k: int) -> np.ndarray:
    """m rows of k distinct columns out of n, each row ascending."""
    if k * k > n:   # rejection sampling would rarely succeed: draw by random permutation keys
        out = np.empty((m, k), dtype=np.int32)
        step = max(1, (1 << 24) // max(n, 1))
        for r0 in range(0, m, step):
            r1 = min(m, r0 + step)
            keys = rng.random((r1 - r0, n))
            out[r0:r1] = np.sort(np.argpartition(keys, k - 1, axis=1)[:, :k], axis=1)
        return out
    cols = np.sort(rng.integers(0, n, size=(m, k), dtype=np.int64), axis=1)
    for _ in range(100):
        bad = np.nonzero((np.diff(cols, axis=1) == 0).any(axis=1))[0]
        if bad.size == 0:
            break
        cols[bad] = np.sort(rng.integers(0, n, size=(bad.size, k), dtype=np.int64), axis=1)
    else:
        raise RuntimeError("could not draw distinct columns")
    return cols.astype(np.int32)


def make_sparse_milp(m: int, n: int, nnz_per_row: int = 10, seed: int = 12345, real_data: bool = False,
                     inf_frac: Tuple[float, float, float] = (0.0, 0.0, 0.0),
                     int_frac: float = 0.5, name: str = "sparse_milp") -> LinearRows:
    """Config C2 / C4 shape: per row ``nnz_per_row`` distinct uniform columns,
    integer coefficients in +-{1..9} (30 % negative), ``int_frac`` Integer variables in
    [0, ub], ub in {1..10}, the rest Continuous; a planted feasible point x* and rows
    ``a.x <= a.x* + s`` (s in {0..3}) with 20-40 % equality rows ``a.x = a.x*``.
    ``inf_frac`` = fractions of continuous variables given (ub=+inf, lb=-inf, both)."""
    rng = np.random.default_rng(seed)
    k = min(nnz_per_row, n)
    col = _distinct_sorted_columns(rng, m, n, k)
    mag = rng.integers(1, 10, size=(m, k)).astype(np.float64)
    if real_data:
        mag = mag + rng.random((m, k))
    sign = np.where(rng.random((m, k)) < 0.3, -1.0, 1.0)
    val = mag * sign
    is_int = rng.random(n) < int_frac
    var_type = np.where(is_int, INTEGER, CONTINUOUS).astype(np.uint8)
    lb = np.zeros(n)
    ub = rng.integers(1, 11, size=n).astype(np.float64)
    if real_data:
        ub = np.where(is_int, ub, ub + rng.random(n))
    xstar = np.where(is_int, np.floor(rng.random(n) * (ub + 1)).clip(0, ub), rng.random(n) * ub)
    if not real_data:
        xstar = np.where(is_int, xstar, np.round(xstar * 4) / 4)  # quarter grid: exact sums
    act = (val * xstar[col]).sum(axis=1)
    slack = rng.integers(0, 4, size=m).astype(np.float64)
    eq_frac = 0.2 + 0.2 * rng.random()
    is_eq = rng.random(m) < eq_frac
    row_ub = np.where(is_eq, act, act + slack)
    row_lb = np.where(is_eq, act, -INF)
    f_ub, f_lb, f_both = inf_frac
    if f_ub + f_lb + f_both > 0:
        u = rng.random(n)
        cont = ~is_int
        ub = np.where(cont & (u < f_ub), INF, ub)
        lb = np.where(cont & (u >= f_ub) & (u < f_ub + f_lb), -INF, lb)
        both = cont & (u >= f_ub + f_lb) & (u < f_ub + f_lb + f_both)
        ub = np.where(both, INF, ub)
        lb = np.where(both, -INF, lb)
    row_ptr = (np.arange(m + 1, dtype=np.int64) * k).astype(np.int32)
    return LinearRows(m=m, n=n, row_ptr=row_ptr, col=col.reshape(-1), val=val.reshape(-1),
                      row_lb=row_lb, row_ub=row_ub, var_type=var_type, lb=lb, ub=ub, name=name,
                      xstar=xstar)


def make_sparse_milp_block(m_block: int, n: int, nnz_per_row: int = 10, seed: int = 777, block: int = 0,
                           int_frac: float = 0.5, name: str = "sparse_milp_block") -> LinearRows:
    """Config C4 shape, one ROW BLOCK of it: the variables (types, box, planted point) come from ``seed``
    alone, so every rank of the row-partitioned mode derives the same n columns, while the block's rows come
    from ``(seed, block)``.  Same distributions as ``make_sparse_milp`` (integer data)."""
    vrng = np.random.default_rng(seed)
    is_int = vrng.random(n) < int_frac
    var_type = np.where(is_int, INTEGER, CONTINUOUS).astype(np.uint8)
    lb = np.zeros(n)
    ub = vrng.integers(1, 11, size=n).astype(np.float64)
    xstar = np.where(is_int, np.floor(vrng.random(n) * (ub + 1)).clip(0, ub), np.round(vrng.random(n) * ub * 4) / 4)
    rng = np.random.default_rng([seed, block + 1])
    k = min(nnz_per_row, n)
    col = _distinct_sorted_columns(rng, m_block, n, k)
    val = rng.integers(1, 10, size=(m_block, k)).astype(np.float64) * np.where(rng.random((m_block, k)) < 0.3, -1.0, 1.0)
    act = (val * xstar[col]).sum(axis=1)
    slack = rng.integers(0, 4, size=m_block).astype(np.float64)
    is_eq = rng.random(m_block) < 0.3
    row_ub = np.where(is_eq, act, act + slack)
    row_lb = np.where(is_eq, act, -INF)
    row_ptr = (np.arange(m_block + 1, dtype=np.int64) * k).astype(np.int32)
    return LinearRows(m=m_block, n=n, row_ptr=row_ptr, col=col.reshape(-1), val=val.reshape(-1), row_lb=row_lb,
                      row_ub=row_ub, var_type=var_type, lb=lb, ub=ub, name=f"{name}[{block}]", xstar=xstar)


def make_knapsack_setcover(m: int = 50_000, n: int = 50_000, nnz_per_row: int = 10, seed: int = 2024,
                           name: str = "knapsack_setcover") -> LinearRows:
    """Config C3 shape: half knapsack rows ``sum w_j x_j <= W`` (w in {1..20}, binaries and
    small integers), half set-cover rows ``sum x_j >= 1`` over binaries; planted point."""
    rng = np.random.default_rng(seed)
    k = min(nnz_per_row, n)
    col = _distinct_sorted_columns(rng, m, n, k)
    is_bin = rng.random(n) < 0.7
    var_type = np.where(is_bin, BINARY, INTEGER).astype(np.uint8)
    lb = np.zeros(n)
    ub = np.where(is_bin, 1.0, rng.integers(2, 6, size=n).astype(np.float64))
    xstar = np.floor(rng.random(n) * (ub + 1)).clip(0, ub)
    mk = m // 2
    val = np.ones((m, k))
    val[:mk] = rng.integers(1, 21, size=(mk, k)).astype(np.float64)
    act = (val * xstar[col]).sum(axis=1)
    row_lb = np.full(m, -INF)
    row_ub = np.full(m, INF)
    row_ub[:mk] = act[:mk] + rng.integers(0, 6, size=mk)
    # set-cover rows: make sure the planted point covers each row
    cover = slice(mk, m)
    need = np.nonzero(act[cover] < 1.0)[0] + mk
    xstar[col[need, 0]] = 1.0
    row_lb[cover] = 1.0
    # re-plant knapsack capacities after the fix-up so x* stays feasible
    act = (val * xstar[col]).sum(axis=1)
    row_ub[:mk] = np.maximum(row_ub[:mk], act[:mk])
    row_ptr = (np.arange(m + 1, dtype=np.int64) * k).astype(np.int32)
    return LinearRows(m=m, n=n, row_ptr=row_ptr, col=col.reshape(-1), val=val.reshape(-1),
                      row_lb=row_lb, row_ub=row_ub, var_type=var_type, lb=lb, ub=ub, name=name,
                      xstar=xstar)


def attach_cutoff(inst: LinearRows, k: int, seed: int, slack: float, box=None) -> LinearRows:
    """Gives ``inst`` an objective cut-off row  c.x <= rhs  (LinearHandler::varBndsFromObj_,
    /root/reference/src/base/LinearHandler.cpp:544-597): k random variables with finite root bounds, random
    coefficients, rhs = minimum activity over ``box`` (default: the root box; tests pass the tightened root box,
    which makes a small slack bite) + slack, so the row moves bounds without making the box trivially infeasible."""
    blb, bub = (inst.lb, inst.ub) if box is None else box
    rng = np.random.default_rng(seed)
    finite = np.nonzero(np.isfinite(inst.lb) & np.isfinite(inst.ub) & (np.abs(inst.lb) < 1e19) & (np.abs(inst.ub) < 1e19))[0]
    cols = np.sort(rng.choice(finite, min(k, len(finite)), replace=False)).astype(np.int32)
    vals = rng.uniform(0.5, 3.0, len(cols)) * rng.choice([-1.0, 1.0], len(cols))
    lo = float(np.sum(np.where(vals > 0, vals * blb[cols], vals * bub[cols])))
    inst.cut_col, inst.cut_val, inst.cut_rhs = cols, vals, lo + slack
    return inst


def attach_binary_objective(inst: LinearRows, n_bin: int, n_other: int, seed: int, slack: float,
                            const: float = 0.0, coef_hi: int = 9) -> LinearRows:
    """Turns ``n_bin`` integer variables whose planted value is 0 or 1 into binaries and gives ``inst`` a linear
    objective over them and ``n_other`` further variables (coefficient +-1), with an incumbent ``slack`` above the
    planted point's objective value.  A binary's coefficient is positive where the planted value is 0 and negative
    where it is 1, so the planted point sits at the binaries' minimum: the setting in which
    NlPresHandler::fixObjBins_ (/root/reference/src/base/NlPresHandler.cpp:1062-1121) fixes the binaries whose
    coefficient alone exceeds the remaining slack, while LinearHandler's cut-off row keeps the planted point.
    ``const`` is the objective's constant term."""
    rng = np.random.default_rng(seed)
    elig = np.nonzero((inst.var_type == INTEGER) & ((inst.xstar == 0) | (inst.xstar == 1)) & (inst.lb <= 0) & (inst.ub >= 1))[0]
    bins = rng.choice(elig, min(n_bin, len(elig)), replace=False)
    rest = np.setdiff1d(np.arange(inst.n), bins)
    others = rng.choice(rest, n_other, replace=False)
    inst.var_type = inst.var_type.copy(); inst.lb = inst.lb.copy(); inst.ub = inst.ub.copy()
    inst.var_type[bins] = BINARY
    inst.lb[bins], inst.ub[bins] = 0.0, 1.0
    coef = np.zeros(inst.n)
    coef[bins] = rng.integers(1, coef_hi, len(bins)) * np.where(inst.xstar[bins] == 0, 1.0, -1.0)
    coef[others] = rng.choice([-1.0, 1.0], len(others))
    cols = np.sort(np.concatenate([bins, others])).astype(np.int32)
    vals = coef[cols]
    inst.cut_col, inst.cut_val = cols, vals
    inst.cut_rhs, inst.obj_const = float(vals @ inst.xstar[cols]) + slack, float(const)
    return inst


def branch_boxes(inst_lb: np.ndarray, inst_ub: np.ndarray, var_type: np.ndarray, n_boxes: int,
                 seed: int = 2024, max_depth: int = 20, continuous_too: bool = False
                 ) -> Tuple[np.ndarray, np.ndarray]:
    """``n_boxes`` node boxes: the root box with d in {1..max_depth} random branching
    perturbations, ``ub := floor(v)`` or ``lb := ceil(v)`` on random integer variables, as
    ``IntVarHandler::getBranches`` creates them (/root/reference/src/base/IntVarHandler.cpp:133-190).
    Returns box-major arrays [n_boxes, n]."""
    rng = np.random.default_rng(seed)
    n = inst_lb.shape[0]
    lbs = np.repeat(inst_lb[None, :], n_boxes, axis=0)
    ubs = np.repeat(inst_ub[None, :], n_boxes, axis=0)
    is_int = (var_type == INTEGER) | (var_type == BINARY)
    cand = np.nonzero((is_int | continuous_too) & np.isfinite(inst_lb) & np.isfinite(inst_ub)
                      & (inst_ub > inst_lb))[0]
    if cand.size == 0:
        return lbs, ubs
    for b in range(n_boxes):
        d = int(rng.integers(1, max_depth + 1))
        for j in rng.choice(cand, size=min(d, cand.size), replace=False):
            lo, hi = lbs[b, j], ubs[b, j]
            if hi - lo < 0.5:
                continue
            v = lo + (hi - lo) * rng.random()
            if is_int[j]:
                if v == np.floor(v):
                    v += 0.5
                if rng.random() < 0.5:
                    ubs[b, j] = np.floor(v)
                else:
                    lbs[b, j] = np.ceil(v)
            else:
                if rng.random() < 0.5:
                    ubs[b, j] = v
                else:
                    lbs[b, j] = v
    return lbs, ubs


def boxes_to_deltas(root_lb: np.ndarray, root_ub: np.ndarray, lbs: np.ndarray, ubs: np.ndarray):
    """Node boxes as sparse branching deltas on the root box, the form in which B&B nodes differ (Node r_mods):
    (delta_ptr [n_boxes+1] int64, delta_var int32, delta_is_upper uint8, delta_val float64)."""
    ptr, var, up, val = [0], [], [], []
    for b in range(lbs.shape[0]):
        jl = np.nonzero(lbs[b] != root_lb)[0]
        ju = np.nonzero(ubs[b] != root_ub)[0]
        var += list(jl) + list(ju)
        up += [0] * len(jl) + [1] * len(ju)
        val += list(lbs[b, jl]) + list(ubs[b, ju])
        ptr.append(len(var))
    return (np.asarray(ptr, np.int64), np.asarray(var, np.int32), np.asarray(up, np.uint8), np.asarray(val, np.float64))


def make_minlp(n: int, n_cons: int, m_lin: int, seed: int = 99, name: str = "minlp"
               ) -> Tuple[LinearRows, Tapes]:
    """Config C5 shape: variables in [l,u] within [-10,10]; 50 % constraints
    ``c_lb <= x_i*x_j + a*x_k <= c_ub`` (tape Var,Var,Mult + linear part), 50 %
    ``x_i^2 + x_j^2 <= r`` as SumList(Sqr,Sqr); plus ``m_lin`` linear rows; planted point."""
    rng = np.random.default_rng(seed)
    lo = 0.0 - rng.integers(0, 11, size=n).astype(np.float64)   # 0.0 - 0 = +0.0 (a -0.0 lower bound trips the
    # reference's BoundsOnRecip: 1/-0.0 = -inf, Operations.cpp:208-211)
    hi = rng.integers(1, 11, size=n).astype(np.float64)
    is_int = rng.random(n) < 0.3
    var_type = np.where(is_int, INTEGER, CONTINUOUS).astype(np.uint8)
    xstar = lo + (hi - lo) * rng.random(n)
    xstar = np.where(is_int, np.round(xstar), xstar).clip(lo, hi)
    cons = []
    for c in range(n_cons):
        i, j, k = rng.choice(n, size=3, replace=False)
        if c % 2 == 0:
            a = float(rng.integers(1, 6)) * (1 if rng.random() < 0.5 else -1)
            v = xstar[i] * xstar[j] + a * xstar[k]
            w = float(rng.integers(0, 4))
            if rng.random() < 0.3:
                clb, cub = v - 0.5 * w, v + 0.5 * w
            else:
                clb, cub = -INF, v + w
            cons.append((Expr.v(int(i)) * Expr.v(int(j)), [(int(k), a)], clb, cub))
        else:
            r = xstar[i] ** 2 + xstar[j] ** 2 + float(rng.integers(0, 10))
            cons.append((Expr.sumlist([Expr.v(int(i)).sqr(), Expr.v(int(j)).sqr()]), [], -INF, r))
    tapes = build_tapes(cons)
    k = min(6, n)
    if m_lin > 0:
        col = _distinct_sorted_columns(rng, m_lin, n, k)
        val = rng.integers(1, 10, size=(m_lin, k)).astype(np.float64) * np.where(rng.random((m_lin, k)) < 0.3, -1.0, 1.0)
        act = (val * xstar[col]).sum(axis=1)
        row_ub = act + rng.integers(0, 4, size=m_lin)
        row_lb = np.where(rng.random(m_lin) < 0.3, act, -INF)
        row_ub = np.where(np.isfinite(row_lb), act, row_ub)
    else:
        col = np.zeros((0, k), np.int32); val = np.zeros((0, k)); row_lb = np.zeros(0); row_ub = np.zeros(0)
    row_ptr = (np.arange(m_lin + 1, dtype=np.int64) * k).astype(np.int32)
    lin = LinearRows(m=m_lin, n=n, row_ptr=row_ptr, col=col.reshape(-1).astype(np.int32),
                     val=val.reshape(-1).astype(np.float64), row_lb=row_lb.astype(np.float64),
                     row_ub=row_ub.astype(np.float64), var_type=var_type, lb=lo, ub=hi, name=name,
                     xstar=xstar)
    return lin, tapes


# --------------------------------------------------------------------------------------
# large-batch forms: node boxes as sparse deltas, vectorised MINLP generator (bench sizes)
# --------------------------------------------------------------------------------------

def branch_deltas(inst_lb: np.ndarray, inst_ub: np.ndarray, var_type: np.ndarray, n_boxes: int, seed: int = 2024,
                  max_depth: int = 20, continuous_too: bool = False):
    """The node boxes of ``branch_boxes`` -- the root box with d in {1..max_depth} random branching perturbations
    (``ub := floor(v)`` or ``lb := ceil(v)`` on random integer variables, IntVarHandler::getBranches,
    /root/reference/src/base/IntVarHandler.cpp:133-190) -- produced directly as sparse deltas, vectorised: the form
    ``mntr_gpu_tighten_nodes`` takes, and the only one that fits in memory at config C5's size (4096 boxes x 1M
    variables).  Own random stream (not box-for-box equal to ``branch_boxes``).  Returns
    (delta_ptr int64 [n_boxes+1], delta_var int32, delta_is_upper uint8, delta_val float64)."""
    rng = np.random.default_rng([seed, 7])
    is_int = (var_type == INTEGER) | (var_type == BINARY)
    cand = np.nonzero((is_int | continuous_too) & np.isfinite(inst_lb) & np.isfinite(inst_ub) & (inst_ub > inst_lb))[0]
    if cand.size == 0:
        return np.zeros(n_boxes + 1, np.int64), np.zeros(0, np.int32), np.zeros(0, np.uint8), np.zeros(0)
    depth = rng.integers(1, max_depth + 1, size=n_boxes)
    box = np.repeat(np.arange(n_boxes, dtype=np.int64), depth)
    var = cand[rng.integers(0, cand.size, size=box.size)].astype(np.int64)
    # one perturbation per (box, variable): drop repeated draws
    key = box * (int(var_type.shape[0]) + 1) + var
    _, first = np.unique(key, return_index=True)
    first.sort()
    box, var = box[first], var[first]
    lo, hi = inst_lb[var], inst_ub[var]
    keep = (hi - lo) >= 0.5
    box, var, lo, hi = box[keep], var[keep], lo[keep], hi[keep]
    v = lo + (hi - lo) * rng.random(box.size)
    vi = is_int[var]
    v = np.where(vi & (v == np.floor(v)), v + 0.5, v)
    up = rng.random(box.size) < 0.5
    val = np.where(vi, np.where(up, np.floor(v), np.ceil(v)), v)
    order = np.argsort(box, kind="stable")
    box, var, up, val = box[order], var[order], up[order], val[order]
    ptr = np.zeros(n_boxes + 1, np.int64)
    np.add.at(ptr, box + 1, 1)
    ptr = np.cumsum(ptr)
    return ptr, var.astype(np.int32), up.astype(np.uint8), val.astype(np.float64)


def deltas_box(root_lb: np.ndarray, root_ub: np.ndarray, deltas, b: int):
    """Dense (lb, ub) of box ``b`` of a delta batch."""
    ptr, var, up, val = deltas
    lb, ub = root_lb.copy(), root_ub.copy()
    for q in range(int(ptr[b]), int(ptr[b + 1])):
        if up[q]:
            ub[var[q]] = val[q]
        else:
            lb[var[q]] = val[q]
    return lb, ub


def slice_deltas(deltas, b0: int, b1: int):
    """Boxes [b0, b1) of a delta batch as a batch of their own."""
    ptr, var, up, val = deltas
    q0, q1 = int(ptr[b0]), int(ptr[b1])
    return (ptr[b0:b1 + 1] - q0).astype(np.int64), var[q0:q1], up[q0:q1], val[q0:q1]


def make_minlp_large(n: int, n_cons: int, m_lin: int, seed: int = 99, name: str = "minlp") -> Tuple[LinearRows, Tapes]:
    """Config C5 at full size: the constraint families of ``make_minlp`` (50 % ``c_lb <= x_i*x_j + a*x_k <= c_ub`` as
    tape Var,Var,Mult + linear part; 50 % ``x_i^2 + x_j^2 <= r`` as SumList(Sqr,Sqr); ``m_lin`` linear rows; planted
    point), with the tapes written directly in the node order ``flatten_expr`` / ``CGraph::finalize`` produce (variable
    nodes by ascending id, operators in post-order) -- vectorised, so 1M constraints take seconds.  Own random stream;
    ``tests/test_instances.py`` checks the tapes against ``build_tapes`` on the same draws."""
    rng = np.random.default_rng([seed, 11])
    # boxes inside [-10, 10] that straddle zero, none with zero as an END POINT at the root: the reference's reverse
    # rule of OpMult divides by an interval that merely touches zero as if zero were excluded (BoundsOnRecip,
    # Operations.cpp:182-212: [0,u] -> [1/u, inf]) and then cuts off feasible points; with a million constraints that
    # proves the ROOT box "infeasible" and every node box with it (measured with make_minlp's [-{0..10}, {1..10}] boxes)
    lo = 0.0 - rng.integers(1, 11, size=n).astype(np.float64)
    hi = rng.integers(1, 11, size=n).astype(np.float64)
    is_int = rng.random(n) < 0.3
    var_type = np.where(is_int, INTEGER, CONTINUOUS).astype(np.uint8)
    xstar = lo + (hi - lo) * rng.random(n)
    xstar = np.where(is_int, np.round(xstar), xstar).clip(lo, hi)
    ijk = rng.integers(0, n, size=(n_cons, 3))
    for _ in range(100):
        bad = (ijk[:, 0] == ijk[:, 1]) | (ijk[:, 0] == ijk[:, 2]) | (ijk[:, 1] == ijk[:, 2])
        if not bad.any():
            break
        ijk[bad] = rng.integers(0, n, size=(int(bad.sum()), 3))
    tapes = minlp_tapes_from_draws(ijk, xstar, rng)
    k = min(6, n)
    if m_lin > 0:
        col = _distinct_sorted_columns(rng, m_lin, n, k)
        val = rng.integers(1, 10, size=(m_lin, k)).astype(np.float64) * np.where(rng.random((m_lin, k)) < 0.3, -1.0, 1.0)
        act = (val * xstar[col]).sum(axis=1)
        row_ub = act + rng.integers(0, 4, size=m_lin)
        row_lb = np.where(rng.random(m_lin) < 0.3, act, -INF)
        row_ub = np.where(np.isfinite(row_lb), act, row_ub)
    else:
        col = np.zeros((0, k), np.int32); val = np.zeros((0, k)); row_lb = np.zeros(0); row_ub = np.zeros(0)
    row_ptr = (np.arange(m_lin + 1, dtype=np.int64) * k).astype(np.int32)
    lin = LinearRows(m=m_lin, n=n, row_ptr=row_ptr, col=col.reshape(-1).astype(np.int32),
                     val=val.reshape(-1).astype(np.float64), row_lb=row_lb.astype(np.float64),
                     row_ub=row_ub.astype(np.float64), var_type=var_type, lb=lo, ub=hi, name=name, xstar=xstar)
    return lin, tapes


def minlp_tapes_from_draws(ijk: np.ndarray, xstar: np.ndarray, rng: np.random.Generator, return_cons: bool = False):
    """Tapes of the two C5 constraint families for the variable triples ``ijk`` (even constraints bilinear + linear
    term, odd ones sum of two squares).  With ``return_cons`` also the same constraints as (Expr, linear, lb, ub)
    tuples for ``build_tapes`` (the check of the direct layout)."""
    n_cons = ijk.shape[0]
    i, j, kk = ijk[:, 0], ijk[:, 1], ijk[:, 2]
    bil = (np.arange(n_cons) % 2) == 0
    a = rng.integers(1, 6, size=n_cons).astype(np.float64) * np.where(rng.random(n_cons) < 0.5, 1.0, -1.0)
    w = rng.integers(0, 4, size=n_cons).astype(np.float64)
    two_sided = rng.random(n_cons) < 0.3
    r_extra = rng.integers(0, 10, size=n_cons).astype(np.float64)
    vb = xstar[i] * xstar[j] + a * xstar[kk]
    c_lb = np.where(bil, np.where(two_sided, vb - 0.5 * w, -INF), -INF)
    c_ub = np.where(bil, np.where(two_sided, vb + 0.5 * w, vb + w), xstar[i] ** 2 + xstar[j] ** 2 + r_extra)
    nodes = np.where(bil, 3, 5)
    tape_ptr = np.concatenate([[0], np.cumsum(nodes)]).astype(np.int64)
    n_nodes = int(tape_ptr[-1])
    op = np.zeros(n_nodes, np.uint8); a0 = np.full(n_nodes, -1, np.int32); a1 = np.full(n_nodes, -1, np.int32)
    base = tape_ptr[:-1]
    vlo, vhi = np.minimum(i, j), np.maximum(i, j)
    idx_i = (i > j).astype(np.int32)            # local index of x_i's variable node (variable nodes ascend by id)
    idx_j = 1 - idx_i
    op[base] = OpVar; a0[base] = vlo
    op[base + 1] = OpVar; a0[base + 1] = vhi
    bb = base[bil]
    op[bb + 2] = OpMult; a0[bb + 2] = idx_i[bil]; a1[bb + 2] = idx_j[bil]
    sb = base[~bil]
    op[sb + 2] = OpSqr; a0[sb + 2] = idx_i[~bil]
    op[sb + 3] = OpSqr; a0[sb + 3] = idx_j[~bil]
    n_sq = int((~bil).sum())
    op[sb + 4] = OpSumList
    a0[sb + 4] = 2 * np.arange(n_sq, dtype=np.int32)
    a1[sb + 4] = 2 * np.arange(n_sq, dtype=np.int32) + 2
    child = np.tile(np.array([2, 3], np.int32), n_sq) if n_sq else np.zeros(1, np.int32)
    lin_cnt = bil.astype(np.int64)
    lin_ptr = np.concatenate([[0], np.cumsum(lin_cnt)]).astype(np.int32)
    lin_col = kk[bil].astype(np.int32); lin_val = a[bil]
    tapes = Tapes(n_cons=n_cons, tape_ptr=tape_ptr.astype(np.int32), op=op, arg0=a0, arg1=a1,
                  cnst=np.zeros(n_nodes), child=child, lin_ptr=lin_ptr,
                  lin_col=lin_col if lin_col.size else np.zeros(1, np.int32),
                  lin_val=lin_val if lin_val.size else np.zeros(1), c_lb=c_lb.astype(np.float64), c_ub=c_ub.astype(np.float64))
    if not return_cons:
        return tapes
    cons = []
    for c in range(n_cons):
        if bil[c]:
            cons.append((Expr.v(int(i[c])) * Expr.v(int(j[c])), [(int(kk[c]), float(a[c]))], float(c_lb[c]), float(c_ub[c])))
        else:
            cons.append((Expr.sumlist([Expr.v(int(i[c])).sqr(), Expr.v(int(j[c])).sqr()]), [], float(c_lb[c]), float(c_ub[c])))
    return tapes, cons


def plant_duplicate_rows(inst: LinearRows, n_dups: int, seed: int) -> LinearRows:
    """A copy of a uniform-row instance in which ``n_dups`` rows are overwritten by a copy, the negative, or a multiple
    (2.5x, -0.5x) of another row, some with shifted bounds: what LinearHandler::dupRows_
    (/root/reference/src/base/LinearHandler.cpp:882-949) looks for."""
    import copy
    m = inst.m
    k = inst.nnz // m
    assert inst.nnz == m * k
    rng = np.random.default_rng([seed, 17])
    col = inst.col.reshape(m, k).copy(); val = inst.val.reshape(m, k).copy()
    rl = inst.row_lb.copy(); ru = inst.row_ub.copy()
    for t in range(n_dups):
        i, j = (int(x) for x in rng.integers(0, m, 2))
        if i == j:
            continue
        col[j] = col[i]
        kind = t % 4
        if kind == 0:
            val[j] = val[i]
        elif kind == 1:
            val[j] = -val[i]; rl[j], ru[j] = -ru[j], -rl[j]
        elif kind == 2:
            val[j] = val[i] * 2.5
        else:
            val[j] = val[i] * (-0.5)
        if np.isfinite(ru[j]):
            ru[j] = ru[j] + float(rng.integers(-2, 3))
    out = copy.copy(inst)
    out.col, out.val, out.row_lb, out.row_ub = col.reshape(-1), val.reshape(-1), rl, ru
    out.name = inst.name + "-dups"
    return out


def make_bigm_instance(n_bin: int, n_cont: int, n_rows: int, seed: int) -> LinearRows:
    """Rows of the kind LinearHandler::coeffImp_ (/root/reference/src/base/LinearHandler.cpp:600-704) improves: variable
    upper / lower bound rows  x - M z <= 0,  x - m z >= 0  with loose big-M coefficients (2-term rows: the implications
    computeImpBounds_ reads, :707-783), one-sided rows mixing continuous variables and binaries with big-M style
    coefficients, and some two-sided rows (never touched).  Variables: binaries first, then continuous in [l, u]."""
    rng = np.random.default_rng([seed, 31])
    n = n_bin + n_cont
    var_type = np.full(n, 4, np.uint8); var_type[:n_bin] = 0
    if n_bin > 4:
        var_type[rng.choice(n_bin, n_bin // 8, replace=False)] = 2          # some ImplBin
    lb = np.zeros(n); ub = np.ones(n)
    lb[n_bin:] = rng.integers(-3, 2, n_cont).astype(np.float64)
    ub[n_bin:] = lb[n_bin:] + rng.integers(1, 12, n_cont)
    if n_bin > 6:
        fixed = rng.choice(n_bin, n_bin // 10, replace=False)               # some binaries fixed (skipped by :637)
        lb[fixed] = ub[fixed] = rng.integers(0, 2, len(fixed))
    rows = []
    # variable bound rows
    for j in range(n_bin, n):
        if rng.random() < 0.6:
            z = int(rng.integers(0, n_bin))
            big = float(ub[j] + rng.integers(0, 6))                          # M >= u, often loose
            rows.append(([(z, -big), (j, 1.0)], -INF, 0.0))                  # x - M z <= 0
        if rng.random() < 0.3:
            z = int(rng.integers(0, n_bin))
            sm = float(rng.integers(0, 3))
            rows.append(([(z, -sm), (j, 1.0)], 0.0, INF))                    # x - m z >= 0
    # mixed one-sided / two-sided rows
    while len(rows) < n_rows:
        k = int(rng.integers(2, 9))
        cols = np.sort(rng.choice(n, k, replace=False))
        terms = []
        for c in cols:
            if c < n_bin:
                a = float(rng.integers(1, 30)) * (1.0 if rng.random() < 0.5 else -1.0)
            else:
                a = float(rng.integers(1, 6)) * (1.0 if rng.random() < 0.6 else -1.0)
            terms.append((int(c), a))
        lo = sum(a * (lb[c] if a > 0 else ub[c]) for c, a in terms)
        hi = sum(a * (ub[c] if a > 0 else lb[c]) for c, a in terms)
        r = rng.random()
        if r < 0.45:
            rows.append((terms, -INF, float(np.floor(lo + (hi - lo) * rng.uniform(0.3, 1.1)))))
        elif r < 0.9:
            rows.append((terms, float(np.ceil(lo + (hi - lo) * rng.uniform(-0.1, 0.7))), INF))
        else:
            mid = lo + (hi - lo) * 0.5
            rows.append((terms, float(np.floor(mid - 2)), float(np.ceil(mid + 2))))
    order = rng.permutation(len(rows))                                       # VUB rows before AND after their users
    rows = [rows[i] for i in order]
    row_ptr = [0]; col = []; val = []; rl = []; ru = []
    for terms, a, b in rows:
        terms = sorted(terms)
        col += [t[0] for t in terms]; val += [t[1] for t in terms]
        row_ptr.append(len(col)); rl.append(a); ru.append(b)
    return LinearRows(m=len(rows), n=n, row_ptr=np.array(row_ptr, np.int32), col=np.array(col, np.int32),
                      val=np.array(val, np.float64), row_lb=np.array(rl), row_ub=np.array(ru), var_type=var_type,
                      lb=lb, ub=ub, name=f"bigm-{seed}")


@dataclass
class QuadRelations:
    """The relations a QuadHandler holds after the reformulation (/root/reference/src/base/QuadHandler.cpp:127-179):
    y = x^2 (x2Funs_: one per x, ascending x) and y = x0 * x1 (x0x1Funs_: ascending (x0, x1), x0 < x1, one per pair)."""
    sq_x: np.ndarray    # int32
    sq_y: np.ndarray
    b_x0: np.ndarray
    b_x1: np.ndarray
    b_y: np.ndarray


def make_quad_relations(n_x: int, n_sq: int, n_bil: int, seed: int):
    """n_x original variables (a mix of continuous, integer and binary), one auxiliary y per relation.  Returns
    (QuadRelations, var_type, lb, ub).  Boxes avoid what makes the REFERENCE assert or divide badly: the auxiliary of a
    square has lb >= 0 (QuadHandler.cpp:1165 asserts it), and factors of a
    product straddle zero or stay away from it (BoundsOnRecip, Operations.cpp:182-212)."""
    rng = np.random.default_rng([seed, 53])
    n = n_x + n_sq + n_bil
    var_type = np.full(n, 4, np.uint8)
    kinds = rng.random(n_x)
    var_type[:n_x][kinds < 0.25] = 1
    var_type[:n_x][kinds < 0.1] = 0
    lb = np.zeros(n); ub = np.zeros(n)
    for j in range(n_x):
        if var_type[j] == 0:
            lb[j], ub[j] = 0.0, 1.0
        else:
            r = rng.random()
            if r < 0.4:
                lb[j] = -float(rng.integers(1, 8)); ub[j] = float(rng.integers(1, 8))          # straddles zero
            elif r < 0.7:
                lb[j] = float(rng.integers(1, 5)); ub[j] = lb[j] + float(rng.integers(1, 8))
            else:
                ub[j] = -float(rng.integers(1, 5)); lb[j] = ub[j] - float(rng.integers(1, 8))
            if var_type[j] == 4:
                lb[j] += 0.25 * float(rng.integers(0, 3)) * (1 if lb[j] > 0 else 0); ub[j] += 0.5 * float(rng.integers(0, 2))
    sq_x = np.sort(rng.choice(n_x, n_sq, replace=False)).astype(np.int32)
    sq_y = (n_x + np.arange(n_sq)).astype(np.int32)
    pairs = set()
    while len(pairs) < n_bil:
        a, b = rng.choice(n_x, 2, replace=False)
        pairs.add((int(min(a, b)), int(max(a, b))))
    pairs = sorted(pairs)
    b_x0 = np.array([p[0] for p in pairs], np.int32); b_x1 = np.array([p[1] for p in pairs], np.int32)
    b_y = (n_x + n_sq + np.arange(n_bil)).astype(np.int32)
    # auxiliaries: some free, some bounded so that the reverse rules bite
    for k, y in enumerate(sq_y):
        x = sq_x[k]
        hi = max(lb[x] ** 2, ub[x] ** 2)
        lb[y] = 0.0 if rng.random() < 0.6 else float(rng.integers(0, 3))        # (never negative: the reference asserts it)
        ub[y] = INF if rng.random() < 0.3 else float(np.ceil(hi * rng.uniform(0.2, 1.2))) + 1.0
    for k, y in enumerate(b_y):
        c = [lb[b_x0[k]] * lb[b_x1[k]], lb[b_x0[k]] * ub[b_x1[k]], ub[b_x0[k]] * lb[b_x1[k]], ub[b_x0[k]] * ub[b_x1[k]]]
        lo, hi = min(c), max(c)
        r = rng.random()
        if r < 0.3:
            lb[y], ub[y] = -INF, INF
        else:
            lb[y] = float(np.floor(lo + (hi - lo) * rng.uniform(0.0, 0.4))) - 0.5
            ub[y] = float(np.ceil(lo + (hi - lo) * rng.uniform(0.6, 1.0))) + 0.5
    return QuadRelations(sq_x, sq_y, b_x0, b_x1, b_y), var_type, lb, ub


def make_quad_relations_planted(n_x: int, n_sq: int, n_bil: int, seed: int):
    """make_quad_relations with a planted point: x* inside the box (integral for integer variables), y* = x*^2 or
    x0* x1*, and every auxiliary's bounds widened to contain y* -- so the ROOT box is consistent and the propagation loop
    of QuadHandler::presolveNode (/root/reference/src/base/QuadHandler.cpp:1214-1239) runs to its fixpoint instead of
    stopping at the first relation.  Returns (QuadRelations, var_type, lb, ub, xstar)."""
    rel, var_type, lb, ub = make_quad_relations(n_x, n_sq, n_bil, seed)
    rng = np.random.default_rng([seed, 59])
    n = len(lb)
    xs = np.zeros(n)
    u = rng.random(n_x)
    xs[:n_x] = lb[:n_x] + u * (ub[:n_x] - lb[:n_x])
    isint = var_type[:n_x] != 4
    xs[:n_x][isint] = np.clip(np.round(xs[:n_x][isint]), np.ceil(lb[:n_x][isint]), np.floor(ub[:n_x][isint]))
    xs[rel.sq_y] = xs[rel.sq_x] ** 2
    xs[rel.b_y] = xs[rel.b_x0] * xs[rel.b_x1]
    for y in np.concatenate([rel.sq_y, rel.b_y]):
        lb[y] = min(lb[y], np.floor(xs[y]) - 0.5)
        ub[y] = max(ub[y], np.ceil(xs[y]) + 0.5)
    lb[rel.sq_y] = np.maximum(lb[rel.sq_y], 0.0)           # the reference asserts y.lb >= 0 for a square (:1376)
    return rel, var_type, lb, ub, xs


def quad_node_boxes(lb, ub, n_x: int, nb: int, seed: int, xstar=None, keep: float = 0.7):
    """nb node boxes on the root box (lb, ub): box 0 is the root; the others move a few original variables' bounds
    inward by one unit (a branching), a fraction ``keep`` of them only where the planted point stays inside (those boxes
    cannot be infeasible), the rest blindly and with a few variables fixed at an end of their range (some of those
    boxes are infeasible)."""
    rng = np.random.default_rng([seed, 61])
    L = np.tile(lb, (nb, 1)); U = np.tile(ub, (nb, 1))
    for b in range(1, nb):
        safe = xstar is not None and rng.random() < keep
        for j in rng.choice(n_x, 6, replace=False):
            if not (np.isfinite(L[b, j]) and np.isfinite(U[b, j])) or U[b, j] - L[b, j] < 2:
                continue
            if L[b, j] < 0 < U[b, j]:
                continue                                   # keep straddling factors straddling (BoundsOnRecip's zero end points)
            if rng.random() < 0.5:
                if not safe or xstar[j] >= L[b, j] + 1.0: L[b, j] += 1.0
            else:
                if not safe or xstar[j] <= U[b, j] - 1.0: U[b, j] -= 1.0
        if not safe:                                       # ... and a few variables fixed at an end of their range
            for j in rng.choice(n_x, 8, replace=False):
                if np.isfinite(L[b, j]) and np.isfinite(U[b, j]) and not (L[b, j] < 0 < U[b, j]):
                    if rng.random() < 0.5: U[b, j] = L[b, j]
                    else: L[b, j] = U[b, j]
    return L, U
