"""Writes tests/golden/tls4_flat.txt: test_instances/tls4.nl as flattened by minotaur_b200/nl_reader.py (taken from
tls4_cases.npz, which tests/test_oracle_golden.py::test_nl_reader_reads_the_reference_instance pins against the .nl
file), in a line format the C++ handler_test can read without ASL:

  n m n_cons
  n lines   : type lb ub
  m lines   : row_lb row_ub k  (col val) x k
  n_cons x  : c_lb c_ub nn klin nchild / nn lines: op arg0 arg1 cnst / klin pairs col val / nchild child indices
  objective : k const / k pairs col val
"""
import os
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
z = np.load(os.path.join(HERE, "tls4_cases.npz"))
P = "tls4_inc."          # the variant that carries the objective (cut_col / cut_val)
rp, col, val = z[P + "row_ptr"], z[P + "col"], z[P + "val"]
n, m = len(z[P + "var_type"]), len(rp) - 1
lbs, ubs = z["tls4.lbs"][0], z["tls4.ubs"][0]
tp = z[P + "t.tape_ptr"]; nc = len(tp) - 1
out = [f"{n} {m} {nc}"]
r = lambda x: repr(float(x))
for j in range(n):
    out.append(f"{int(z[P + 'var_type'][j])} {r(lbs[j])} {r(ubs[j])}")
for i in range(m):
    terms = " ".join(f"{int(col[t])} {r(val[t])}" for t in range(rp[i], rp[i + 1]))
    out.append(f"{r(z[P + 'row_lb'][i])} {r(z[P + 'row_ub'][i])} {rp[i + 1] - rp[i]} {terms}")
op, a0, a1, cn = z[P + "t.op"], z[P + "t.arg0"], z[P + "t.arg1"], z[P + "t.cnst"]
child, lp, lc, lv = z[P + "t.child"], z[P + "t.lin_ptr"], z[P + "t.lin_col"], z[P + "t.lin_val"]
OpSumList = 30
for c in range(nc):
    b, e = tp[c], tp[c + 1]
    kids = []
    rows = []
    for i in range(b, e):
        if op[i] == OpSumList:
            x0 = len(kids); kids += [int(k) for k in child[a0[i]:a1[i]]]; x1 = len(kids)
            rows.append(f"{int(op[i])} {x0} {x1} {r(cn[i])}")
        else:
            rows.append(f"{int(op[i])} {int(a0[i])} {int(a1[i])} {r(cn[i])}")
    klin = lp[c + 1] - lp[c]
    out.append(f"{r(z[P + 't.c_lb'][c])} {r(z[P + 't.c_ub'][c])} {e - b} {klin} {len(kids)}")
    out += rows
    out.append(" ".join(f"{int(lc[q])} {r(lv[q])}" for q in range(lp[c], lp[c + 1])) or "-")
    out.append(" ".join(str(k) for k in kids) or "-")
cc, cv = z[P + "cut_col"], z[P + "cut_val"]
out.append(f"{len(cc)} 0.0")
out.append(" ".join(f"{int(cc[t])} {r(cv[t])}" for t in range(len(cc))))
open(os.path.join(HERE, "tls4_flat.txt"), "w").write("\n".join(out) + "\n")
print("wrote tls4_flat.txt:", n, "variables,", m, "rows,", nc, "CGraph constraints,", len(cc), "objective terms")
