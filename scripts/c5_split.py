"""C5-shaped batch: time of the linear handler alone, the nonlinear handler alone, and both (one presolveNode pass)."""
import os, sys, numpy as np, torch
sys.path.insert(0, '/root/repo')
from minotaur_b200 import engine as E
from minotaur_b200.instances import make_minlp, branch_boxes
ncons, nb = 200_000, 1024
lin, tapes = make_minlp(n=ncons, n_cons=ncons, m_lin=ncons // 10, seed=99)
dev = torch.device('cuda', 0)
lbs, ubs = branch_boxes(lin.lb, lin.ub, lin.var_type, nb, seed=99, max_depth=10, continuous_too=True)
eng = E.GpuBoundEngine(0); eng.load_linear(lin); eng.load_cgraph(tapes)
ld = eng.box_ld(nb)
boxes = torch.empty((lin.n, ld, 2), dtype=torch.float64, device=dev)
verdict = torch.zeros(ld, dtype=torch.int32, device=dev); rounds = torch.zeros(ld, dtype=torch.int32, device=dev); nnz = torch.zeros(ld, dtype=torch.int64, device=dev)
eng.boxes_upload(lbs, ubs, boxes.data_ptr()); pristine = boxes.clone()
for name, h in (("linear", E.HANDLERS_LINEAR), ("nonlinear", E.HANDLERS_NONLINEAR), ("both", E.HANDLERS_ALL)):
    ms = []
    for rep in range(3):
        boxes.copy_(pristine); torch.cuda.synchronize()
        st = eng.tighten_dev(nb, boxes.data_ptr(), verdict.data_ptr(), rounds.data_ptr(), nnz.data_ptr(), loop=E.LOOP_SIMPLEPRESOLVE, handlers=h)
        ms.append(st.kernel_ms)
    print(f"C5 {name}: {min(ms[1:]):.2f} ms rounds hist {np.bincount(rounds[:nb].cpu().numpy())[:6]}", flush=True)
