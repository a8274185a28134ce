"""K3 timing probe on the C3 workload (8192 boxes, 50k-row knapsack/set-cover): kernel ms of the fixpoint call."""
import sys, numpy as np, torch
sys.path.insert(0, '/root/repo')
from minotaur_b200 import engine as E
from minotaur_b200.instances import make_knapsack_setcover, branch_boxes
inst = make_knapsack_setcover(50_000, 50_000, 10, seed=2024)
nb = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
lbs, ubs = branch_boxes(inst.lb, inst.ub, inst.var_type, nb, seed=2024, max_depth=20)
eng = E.GpuBoundEngine(0); eng.load_linear(inst)
dev = torch.device('cuda', 0)
ld = eng.box_ld(nb)
boxes = torch.empty((inst.n, ld, 2), dtype=torch.float64, device=dev)
verdict = torch.zeros(ld, dtype=torch.int32, device=dev); rounds = torch.zeros(ld, dtype=torch.int32, device=dev); nnz = torch.zeros(ld, dtype=torch.int64, device=dev)
eng.boxes_upload(lbs, ubs, boxes.data_ptr()); pristine = boxes.clone()
ms = []
for rep in range(4):
    boxes.copy_(pristine); torch.cuda.synchronize()
    st = eng.tighten_dev(nb, boxes.data_ptr(), verdict.data_ptr(), rounds.data_ptr(), nnz.data_ptr())
    ms.append(st.kernel_ms)
r = rounds[:nb].cpu().numpy()
print(f"K3 C3 nb={nb}: {min(ms[1:]):.2f} ms (runs {['%.2f' % m for m in ms]})  nnz={int(nnz[:nb].sum())/1e9:.3f}G  "
      f"infeasible={int((verdict[:nb] != 0).sum())}  rounds hist={np.bincount(r)[:8]}  checksum={float(boxes.sum()):.6f}")
