import sys, time, numpy as np, torch
sys.path.insert(0, '/root/repo')
from minotaur_b200 import engine as E
from minotaur_b200.instances import *
inst = make_knapsack_setcover(50_000, 50_000, 10, seed=2024)
nb = 8192
lbs, ubs = branch_boxes(inst.lb, inst.ub, inst.var_type, nb, seed=2024, max_depth=20)
eng = E.GpuBoundEngine(0); eng.load_linear(inst)
dev = torch.device('cuda', 0)
ld = eng.box_ld(nb)
boxes = torch.empty((inst.n, ld, 2), dtype=torch.float64, device=dev)
verdict = torch.zeros(ld, dtype=torch.int32, device=dev); rounds = torch.zeros(ld, dtype=torch.int32, device=dev); nnz = torch.zeros(ld, dtype=torch.int64, device=dev)
eng.boxes_upload(lbs, ubs, boxes.data_ptr()); pristine = boxes.clone()
for mr in (1, 2, 3, 4, 6, 0):
    for rep in range(2):
        boxes.copy_(pristine); torch.cuda.synchronize()
        st = eng.tighten_dev(nb, boxes.data_ptr(), verdict.data_ptr(), rounds.data_ptr(), nnz.data_ptr(), max_rounds=mr)
    r = rounds[:nb].cpu().numpy()
    print(f"max_rounds={mr}: {st.kernel_ms:.2f} ms  nnz={int(nnz[:nb].sum())/1e9:.3f}G  rounds hist={np.bincount(r)[:12]}")
