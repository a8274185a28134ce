"""Free-format MPS reader -> the flat linear rows that cross the C ABI.

It mirrors ``Reader::readMps`` of the reference (/root/reference/src/base/Reader.cpp:42-473) token for token, its
oddities included, so that a problem read here is the problem the reference's ``bnb`` would presolve:

* a line is split on white space; a line whose first token starts with ``*`` is a comment; a section keyword counts
  only when it starts in column 1 (``NAME ROWS COLUMNS RHS RANGES BOUNDS ENDATA``);
* ROWS: ``N G L E``; the first ``N`` row becomes the objective (Minimize, constant = -rhs), later ones are ignored;
* COLUMNS: variables are numbered by first appearance with the box ``[0, +inf)``; ``'MARKER' 'INTORG'`` ..
  ``'INTEND'`` makes them Integer; a coefficient seen twice for the same (row, column) is ADDED (``incTerm``,
  LinearFunction.cpp:133-142) and an entry with ``|a| <= 1e-9`` after the addition is dropped;
* RHS / RANGES: only the first set name seen is used, other sets are ignored; ranges follow Reader.cpp:404-431 --
  ``G``: ``[rhs, rhs+|R|]``, ``L``: ``[rhs-|R|, rhs]``, ``E``: ``R>0`` ``[rhs, rhs+R]``, and ``R<=0`` gives
  ``lb = rhs - R`` (the reference's sign: ``lb > ub`` for ``R < 0``), reproduced as is;
* BOUNDS (:337-391): ``LO UP FX FR MI PL BV LI UI``; ``UP`` with a negative value on a variable whose lower bound is
  still 0 also sets the lower bound to -inf; ``BV`` only changes the type (the box stays ``[0, +inf)`` unless other
  lines set it); the column of the FIRST bounds line is not checked against the COLUMNS section (the reference's
  else-if chain), so an unknown column there raises here like it crashes there.

The terms of a row come out in ascending variable index (``VariableGroup`` is ordered by variable id, Types.h:496).
"""
from __future__ import annotations

from typing import Dict, List, Tuple

import numpy as np

from .instances import BINARY, CONTINUOUS, INTEGER, INF, LinearRows

_SECTIONS = {"NAME": 1, "ROWS": 2, "COLUMNS": 3, "RHS": 4, "RANGES": 5, "BOUNDS": 6, "ENDATA": 7}
_DROP = 1e-9       # LinearFunction::tol_


class MpsError(ValueError):
    pass


def _inc(terms: Dict[int, float], j: int, a: float) -> None:
    """LinearFunction::incTerm (LinearFunction.cpp:133-142): add, and drop the term if it cancels."""
    if abs(a) <= _DROP:
        return
    if j in terms:
        s = terms[j] + a
        if abs(s) < _DROP:
            del terms[j]
        else:
            terms[j] = s
    else:
        terms[j] = a


def read_mps(path: str) -> LinearRows:
    rowtypes: List[str] = []
    rowrhs: List[float] = []
    rowrange: List[float] = []
    rownames: Dict[str, int] = {}
    rowterms: List[Dict[int, float]] = []
    colnames: Dict[str, int] = {}
    lb: List[float] = []
    ub: List[float] = []
    vtype: List[int] = []
    cur_type = CONTINUOUS
    rhsid = rangeid = bndid = ""
    section = 0
    with open(path) as f:
        for lcnt, line in enumerate(f, 1):
            if section == 7:
                break
            w = line.split()
            if not w or w[0][0] == "*":
                continue
            if line[0] == w[0][0] and w[0] in _SECTIONS:
                section = _SECTIONS[w[0]]
                continue
            if section == 0:
                raise MpsError(f"line {lcnt}: data before any section")
            if section == 1:
                continue
            if section == 2:
                if w[0][0] not in "NGLE":
                    raise MpsError(f"line {lcnt}: unexpected word {w[0]}")
                if len(w) != 2:
                    raise MpsError(f"line {lcnt}: a ROWS line has a type and a name")
                if w[1] in rownames:
                    raise MpsError(f"line {lcnt}: row {w[1]} seen more than once")
                rownames[w[1]] = len(rowtypes)
                rowtypes.append(w[0][0]); rowrhs.append(INF); rowrange.append(INF); rowterms.append({})
            elif section == 3:
                if len(w) < 3:
                    raise MpsError(f"line {lcnt}: not enough fields in COLUMNS")
                if w[1] == "'MARKER'":
                    if w[2] == "'INTORG'":
                        if cur_type == INTEGER:
                            raise MpsError(f"line {lcnt}: 'INTORG' within 'INTORG'")
                        cur_type = INTEGER
                    elif w[2] == "'INTEND'":
                        if cur_type == CONTINUOUS:
                            raise MpsError(f"line {lcnt}: 'INTEND' outside 'INTORG'")
                        cur_type = CONTINUOUS
                    else:
                        raise MpsError(f"line {lcnt}: unknown marker {w[2]}")
                    continue
                if w[1] not in rownames:
                    continue                       # (the reference logs the undeclared row and goes on)
                if w[0] not in colnames:
                    colnames[w[0]] = len(lb)
                    lb.append(0.0); ub.append(INF); vtype.append(cur_type)
                j = colnames[w[0]]
                _inc(rowterms[rownames[w[1]]], j, float(w[2]))
                if len(w) > 3:
                    if len(w) < 5:
                        raise MpsError(f"line {lcnt}: not enough fields in COLUMNS")
                    if w[3] not in rownames:
                        raise MpsError(f"line {lcnt}: row {w[3]} undeclared")
                    _inc(rowterms[rownames[w[3]]], j, float(w[4]))
            elif section in (4, 5):
                target = rowrhs if section == 4 else rowrange
                sid = rhsid if section == 4 else rangeid
                if sid == "":
                    sid = w[0]
                    if section == 4:
                        rhsid = sid
                    else:
                        rangeid = sid
                elif w[0] != sid:
                    continue                       # another set: ignored
                if len(w) < 3:
                    raise MpsError(f"line {lcnt}: not enough fields")
                if w[1] not in rownames:
                    continue
                target[rownames[w[1]]] = float(w[2])
                if len(w) > 3:
                    if len(w) < 5:
                        raise MpsError(f"line {lcnt}: not enough fields")
                    if w[3] not in rownames:
                        raise MpsError(f"line {lcnt}: row {w[3]} undeclared")
                    target[rownames[w[3]]] = float(w[4])
            elif section == 6:
                if len(w) < 3:
                    raise MpsError(f"line {lcnt}: not enough fields in BOUNDS")
                if bndid == "":
                    bndid = w[1]
                elif w[1] != bndid:
                    continue
                if w[2] not in colnames:
                    raise MpsError(f"line {lcnt}: column {w[2]} undeclared")
                key = w[0]
                dval = float(w[3]) if len(w) > 3 else INF
                j = colnames[w[2]]
                if key == "LO":
                    lb[j] = dval
                elif key == "UP":
                    if dval < 0.0 and lb[j] == 0.0:
                        lb[j] = -INF
                    ub[j] = dval
                elif key == "FX":
                    lb[j] = ub[j] = dval
                elif key == "FR":
                    lb[j], ub[j] = -INF, INF
                elif key == "MI":
                    lb[j] = -INF
                elif key == "PL":
                    ub[j] = INF
                elif key == "BV":
                    vtype[j] = BINARY
                elif key == "LI":
                    vtype[j] = INTEGER; lb[j] = dval
                elif key == "UI":
                    vtype[j] = INTEGER; ub[j] = dval
                else:
                    raise MpsError(f"line {lcnt}: unknown bound type {key}")

    # constraints in ROWS order; the first N row is the objective
    row_ptr = [0]
    col: List[int] = []
    val: List[float] = []
    rlb: List[float] = []
    rub: List[float] = []
    obj: Tuple[np.ndarray, np.ndarray, float] | None = None
    for i, t in enumerate(rowtypes):
        rhs = 0.0 if rowrhs[i] == INF else rowrhs[i]
        rng = rowrange[i]
        items = sorted(rowterms[i].items())
        if t == "N":
            if obj is None:
                obj = (np.array([j for j, _ in items], np.int32), np.array([a for _, a in items], np.float64),
                       0.0 if rowrhs[i] == INF else -rowrhs[i])
            continue
        if t == "G":
            lo, hi = rhs, (rhs + abs(rng) if rng < INF else INF)
        elif t == "L":
            hi = rhs
            lo = rhs - abs(rng) if rng < INF else -INF
        else:
            if rng == INF:
                lo = hi = rhs
            elif rng > 0.0:
                lo, hi = rhs, rhs + rng
            else:
                hi = rhs
                lo = hi - rng
        for j, a in items:
            col.append(j); val.append(a)
        row_ptr.append(len(col)); rlb.append(lo); rub.append(hi)
    n, m = len(lb), len(rlb)
    out = LinearRows(m=m, n=n, row_ptr=np.array(row_ptr, np.int32), col=np.array(col, np.int32),
                     val=np.array(val, np.float64), row_lb=np.array(rlb, np.float64), row_ub=np.array(rub, np.float64),
                     var_type=np.array(vtype, np.uint8), lb=np.array(lb, np.float64), ub=np.array(ub, np.float64),
                     name=path)
    if obj is not None:
        out.cut_col, out.cut_val, out.obj_const = obj
    return out
