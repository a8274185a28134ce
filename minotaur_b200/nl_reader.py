"""Minimal reader for AMPL text-format ('g') .nl files -> the flat description that crosses the C ABI.

No ASL: the segments ``b r C O x k J G`` are parsed directly, and the problem is assembled the way the
reference does it in ``AMPLInterface::copyInstanceFromASL2_``
(/root/reference/src/interfaces/AMPLInterface.cpp:675-782):

* variable order and typing follow ``addVariablesFromASL_`` (:420-565): the first ``max(nlvc, nlvo)`` variables
  are the nonlinear ones (continuous, then integer, inside each of the groups both / constraints-only /
  objective-only), then the linear continuous ones, ``nbv`` Binary, ``niv`` Integer;
* the first ``nlc`` constraints are the nonlinear ones: expression -> CGraph (``getCGraph_`` :986-1165, the
  children of a binary operator in file order), linear part from the ``J`` segment; the remaining
  constraints are linear rows;
* a linear objective gives the (col, val, constant) that ``LinearHandler::varBndsFromObj_`` uses as cut-off row.

Only what the hot path consumes is kept.  Unsupported: binary format, network constraints, defined variables,
complementarities, SOS suffixes, nonlinear objectives (their linear part is still returned).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Tuple

import numpy as np

from .instances import (BINARY, CONTINUOUS, INTEGER, INF, Expr, LinearRows, OpAbs, OpCeil, OpCos, OpDiv, OpExp, OpFloor,
                        OpLog, OpLog10, OpMinus, OpMult, OpPlus, OpPowK, OpSin, OpSqr, OpSqrt, OpSumList, OpUMinus,
                        Tapes, build_tapes)

# AMPL opcodes (opcode.hd) -> (Minotaur OpCode, arity); mirrors AMPLInterface::getCGraph_
_BINARY = {0: OpPlus, 1: OpMinus, 2: OpMult, 3: OpDiv}
_UNARY = {13: OpFloor, 14: OpCeil, 15: OpAbs, 16: OpUMinus, 39: OpSqrt, 41: OpSin, 42: OpLog10, 43: OpLog, 44: OpExp,
          46: OpCos}


@dataclass
class NlProblem:
    lin: LinearRows            # the linear rows (constraints nlc..n_con-1), variable types, root box, objective
    tapes: Tapes               # the nonlinear constraints 0..nlc-1 (n_cons may be 0)
    n_var: int
    n_con: int
    obj_sense_max: bool
    obj_nonlinear: bool
    x0: np.ndarray             # primal starting point of the x segment (zeros where absent)


def _parse_expr(lines: List[str], pos: int) -> Tuple[Expr, int]:
    """Prefix expression starting at lines[pos]; returns (Expr, next position)."""
    tok = lines[pos].split("#")[0].strip()
    pos += 1
    kind, rest = tok[0], tok[1:]
    if kind == "n":
        return Expr.c(float(rest)), pos
    if kind == "v":
        return Expr.v(int(rest)), pos
    if kind != "o":
        raise ValueError(f".nl: unexpected expression token {tok!r}")
    code = int(rest)
    if code in _BINARY:
        a, pos = _parse_expr(lines, pos)
        b, pos = _parse_expr(lines, pos)
        return Expr(_BINARY[code], (a, b)), pos
    if code in _UNARY:
        a, pos = _parse_expr(lines, pos)
        return Expr(_UNARY[code], (a,)), pos
    if code == 54:                       # OPSUMLIST: operand count on the next line
        k = int(lines[pos].split("#")[0]); pos += 1
        kids = []
        for _ in range(k):
            e, pos = _parse_expr(lines, pos)
            kids.append(e)
        return Expr(OpSumList, tuple(kids)), pos
    if code == 5:                        # OPPOW with a constant exponent -> OP1POW / OP2POW as ASL rewrites them
        a, pos = _parse_expr(lines, pos)
        b, pos = _parse_expr(lines, pos)
        if b.op in (21, 14) and b.value == 2.0:
            return Expr(OpSqr, (a,)), pos
        if b.op in (21, 14):
            return Expr(OpPowK, (a, b)), pos
        raise ValueError(".nl: general power expr^expr is not supported")
    raise ValueError(f".nl: opcode o{code} is not supported")


def read_nl(path: str) -> NlProblem:
    with open(path, "r") as f:
        lines = f.read().splitlines()
    if not lines or not lines[0].startswith("g"):
        raise ValueError(".nl: only the text ('g') format is supported")

    def ints(k):
        return [int(float(t)) for t in lines[k].split("#")[0].split()]

    n_var, n_con, n_obj = ints(1)[:3]
    nlc, nlo = ints(2)[:2]
    if any(ints(3)[:2]):
        raise ValueError(".nl: network constraints are not supported")
    nlvc, nlvo, nlvb = ints(4)[:3]
    nwv = ints(5)[0]
    nbv, niv, nlvbi, nlvci, nlvoi = ints(6)[:5]
    if nwv:
        raise ValueError(".nl: linear arcs are not supported")
    if any(ints(9)[:5]):
        raise ValueError(".nl: defined variables (common expressions) are not supported")

    # ---- variable types, addVariablesFromASL_ ----
    vt = np.full(n_var, CONTINUOUS, np.uint8)
    i = nlvb - nlvbi
    vt[i:i + nlvbi] = INTEGER; i += nlvbi
    i += nlvc - (nlvb + nlvci)
    vt[i:i + nlvci] = INTEGER; i += nlvci
    if nlvo > nlvc:
        i += nlvo - (nlvc + nlvoi)
        vt[i:i + nlvoi] = INTEGER; i += nlvoi
    first_bin = n_var - (niv + nbv)
    vt[first_bin:first_bin + nbv] = BINARY
    vt[first_bin + nbv:] = INTEGER

    lb = np.full(n_var, -INF); ub = np.full(n_var, INF)
    r_lb = np.full(n_con, -INF); r_ub = np.full(n_con, INF)
    x0 = np.zeros(n_var)
    con_expr = [None] * n_con
    con_lin: List[List[Tuple[int, float]]] = [[] for _ in range(n_con)]
    obj_lin: List[Tuple[int, float]] = []
    obj_const, obj_nonlinear, obj_max = 0.0, False, False

    def bounds(pos, count, lo, hi):
        for k in range(count):
            t = lines[pos + k].split("#")[0].split()
            code = int(t[0])
            if code == 0: lo[k], hi[k] = float(t[1]), float(t[2])
            elif code == 1: hi[k] = float(t[1])
            elif code == 2: lo[k] = float(t[1])
            elif code == 3: pass
            elif code == 4: lo[k] = hi[k] = float(t[1])
            else: raise ValueError(".nl: complementarity constraints are not supported")
        return pos + count

    pos = 10
    while pos < len(lines):
        head = lines[pos].split("#")[0].strip()
        if not head:
            pos += 1
            continue
        kind = head[0]
        if kind == "C":
            idx = int(head[1:])
            con_expr[idx], pos = _parse_expr(lines, pos + 1)
        elif kind == "O":
            parts = head[1:].split()
            obj_max = len(parts) > 1 and int(parts[1]) == 1
            e, pos = _parse_expr(lines, pos + 1)
            if int(parts[0]) == 0:
                if e.op in (21, 14):                 # OpNum / OpInt: the objective constant
                    obj_const = e.value
                else:
                    obj_nonlinear = True
        elif kind == "b":
            pos = bounds(pos + 1, n_var, lb, ub)
        elif kind == "r":
            pos = bounds(pos + 1, n_con, r_lb, r_ub)
        elif kind == "x":
            k = int(head[1:])
            for q in range(k):
                t = lines[pos + 1 + q].split("#")[0].split()
                x0[int(t[0])] = float(t[1])
            pos += 1 + k
        elif kind == "k":
            pos += 1 + int(head[1:])
        elif kind == "J":
            idx, k = (int(t) for t in head[1:].split())
            for q in range(k):
                t = lines[pos + 1 + q].split("#")[0].split()
                con_lin[idx].append((int(t[0]), float(t[1])))
            pos += 1 + k
        elif kind == "G":
            idx, k = (int(t) for t in head[1:].split())
            for q in range(k):
                t = lines[pos + 1 + q].split("#")[0].split()
                if idx == 0:
                    obj_lin.append((int(t[0]), float(t[1])))
            pos += 1 + k
        elif kind == "d":
            pos += 1 + int(head[1:])
        else:
            raise ValueError(f".nl: segment {head!r} is not supported")

    # ---- nonlinear constraints 0..nlc-1 -> tapes ----
    cons = []
    for c in range(nlc):
        e = con_expr[c]
        if e is None or e.op in (21, 14):
            raise ValueError(f".nl: constraint {c} is declared nonlinear but has no expression")
        cons.append((e, [(j, a) for j, a in con_lin[c] if a != 0.0], r_lb[c], r_ub[c]))
    tapes = build_tapes(cons) if cons else Tapes(
        n_cons=0, tape_ptr=np.zeros(1, np.int32), op=np.zeros(0, np.uint8), arg0=np.zeros(0, np.int32),
        arg1=np.zeros(0, np.int32), cnst=np.zeros(0), child=np.zeros(1, np.int32), lin_ptr=np.zeros(1, np.int32),
        lin_col=np.zeros(1, np.int32), lin_val=np.zeros(1), c_lb=np.zeros(0), c_ub=np.zeros(0))

    # ---- linear constraints nlc..n_con-1 -> CSR (columns ascending, zero coefficients of the J segment dropped
    #      as LinearFunction::addTerm drops them) ----
    rp, col, val = [0], [], []
    for c in range(nlc, n_con):
        for j, a in sorted(con_lin[c]):
            if abs(a) > 1e-9:
                col.append(j); val.append(a)
        rp.append(len(col))
    m = n_con - nlc
    lin = LinearRows(m=m, n=n_var, row_ptr=np.asarray(rp, np.int32), col=np.asarray(col, np.int32),
                     val=np.asarray(val, np.float64), row_lb=r_lb[nlc:].copy(), row_ub=r_ub[nlc:].copy(),
                     var_type=vt, lb=lb, ub=ub, name=path.rsplit("/", 1)[-1])
    if obj_lin:
        ol = [(j, a) for j, a in sorted(obj_lin) if abs(a) > 1e-9]
        lin.cut_col = np.asarray([j for j, _ in ol], np.int32)
        lin.cut_val = np.asarray([a for _, a in ol], np.float64)
        lin.cut_rhs = INF                      # no incumbent yet
        lin.obj_const = obj_const
    return NlProblem(lin=lin, tapes=tapes, n_var=n_var, n_con=n_con, obj_sense_max=obj_max,
                     obj_nonlinear=obj_nonlinear, x0=x0)
