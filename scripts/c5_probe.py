"""K3+K4 timing probe on the C5-shaped MINLP batch (one presolveNode pass), for several box counts / cluster sizes."""
import os, sys, numpy as np, torch
sys.path.insert(0, '/root/repo')
from minotaur_b200 import engine as E
from minotaur_b200.instances import make_minlp, branch_boxes
ncons = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000
lin, tapes = make_minlp(n=ncons, n_cons=ncons, m_lin=ncons // 10, seed=99)
dev = torch.device('cuda', 0)
for nb in (256, 512, 1024, 2048):
    lbs, ubs = branch_boxes(lin.lb, lin.ub, lin.var_type, nb, seed=99, max_depth=10, continuous_too=True)
    for cl in ("auto", "1", "2", "4", "8"):
        if cl == "auto": os.environ.pop("MNTR_GPU_CLUSTER", None)
        else: os.environ["MNTR_GPU_CLUSTER"] = cl
        eng = E.GpuBoundEngine(0); eng.load_linear(lin); eng.load_cgraph(tapes)
        ld = eng.box_ld(nb)
        boxes = torch.empty((lin.n, ld, 2), dtype=torch.float64, device=dev)
        verdict = torch.zeros(ld, dtype=torch.int32, device=dev); rounds = torch.zeros(ld, dtype=torch.int32, device=dev); nnz = torch.zeros(ld, dtype=torch.int64, device=dev)
        eng.boxes_upload(lbs, ubs, boxes.data_ptr()); pristine = boxes.clone()
        ms = []
        for rep in range(3):
            boxes.copy_(pristine); torch.cuda.synchronize()
            st = eng.tighten_dev(nb, boxes.data_ptr(), verdict.data_ptr(), rounds.data_ptr(), nnz.data_ptr(), loop=E.LOOP_SIMPLEPRESOLVE)
            ms.append(st.kernel_ms)
        print(f"C5 cons={ncons} nb={nb} cluster={cl}: {min(ms[1:]):.2f} ms  infeasible={int((verdict[:nb] != 0).sum())}", flush=True)
        eng.close()
