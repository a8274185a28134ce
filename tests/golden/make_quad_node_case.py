"""Writes tests/golden/quad_node_case.txt: a small problem of relations y = x^2, y = x0 x1 with a planted point and a few
node boxes, read by oracle/_ref/quad_patch_test (the reference's QuadHandler with handler/quad_handler_gpu.patch applied,
run on the device, against the same handler run on the host).  Plain text: one array per line."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from minotaur_b200.instances import make_quad_relations_planted, quad_node_boxes  # noqa: E402

rel, vt, lb, ub, xs = make_quad_relations_planted(120, 40, 90, 17)
L, U = quad_node_boxes(lb, ub, 120, 16, 17, xs)


def row(a, fmt):
    return " ".join(fmt % v for v in a) + "\n"


with open(os.path.join(ROOT, "tests", "golden", "quad_node_case.txt"), "w") as f:
    n = len(lb)
    f.write(f"{n} {len(rel.sq_x)} {len(rel.b_x0)} {len(L)}\n")
    f.write(row(vt, "%d"))
    f.write(row(lb, "%.17g")); f.write(row(ub, "%.17g"))
    for a in (rel.sq_x, rel.sq_y, rel.b_x0, rel.b_x1, rel.b_y):
        f.write(row(a, "%d"))
    for b in range(len(L)):
        f.write(row(L[b], "%.17g")); f.write(row(U[b], "%.17g"))
print("written", n, "variables,", len(rel.sq_x) + len(rel.b_x0), "relations,", len(L), "boxes")
