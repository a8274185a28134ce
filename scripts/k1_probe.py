"""K1 timing probe on C2 (single box, Jacobi fixpoint in one launch): kernel us with / without an L2 flush before
the launch; with MNTR_GPU_TRACE=1 the library prints the in-kernel phase trace of every call to stderr."""
import os, sys, numpy as np, torch
sys.path.insert(0, '/root/repo')
from minotaur_b200 import engine as E
if os.environ.get("MNTR_GPU_LIB"):          # dev: a variant build (scripts/build_variant.sh)
    E.LIB_PATH = os.path.abspath(os.environ["MNTR_GPU_LIB"])
from minotaur_b200.instances import make_sparse_milp
m = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
flags = int(sys.argv[2]) if len(sys.argv) > 2 else 0     # 1: per-round kernels (K5), 2: staged rows
inst = make_sparse_milp(m, m, 10, seed=12345)
eng = E.GpuBoundEngine(0); eng.load_linear(inst)
dev = torch.device('cuda', 0)
stream = torch.cuda.ExternalStream(eng.stream_handle(), device=dev)
root_lb = torch.from_numpy(inst.lb).to(dev); root_ub = torch.from_numpy(inst.ub).to(dev)
w_lb = torch.empty_like(root_lb); w_ub = torch.empty_like(root_ub)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for do_flush in (True, False):
    us = []
    for it in range(12):
        with torch.cuda.stream(stream):
            w_lb.copy_(root_lb); w_ub.copy_(root_ub)
            if do_flush: flush.zero_()
        v, r, z = eng.tighten_single_dev(w_lb.data_ptr(), w_ub.data_ptr(), flags=flags)
        st = eng.stats(); us.append(st.kernel_ms * 1e3)
    us = np.array(us[4:])
    print(f"K1 {os.path.basename(E.LIB_PATH)} flags={flags} m={m} flush={do_flush}: {us.mean():.1f} us (min {us.min():.1f})  rounds={r} nnz={z} verdict={v} "
          f"checksum={float(w_lb.sum() + w_ub.sum()):.6f} rows_ms={st.rows_ms:.3f} vars_ms={st.vars_ms:.3f}", flush=True)
