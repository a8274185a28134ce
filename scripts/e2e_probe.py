"""Split of the end-to-end C-ABI call (mntr_gpu_tighten with pinned host buffers) on C2.
The device span of the call is only recorded with MNTR_GPU_TIMING=1 (the event records cost about 7 us)."""
import sys, time, numpy as np, torch
sys.path.insert(0, '/root/repo')
from minotaur_b200 import engine as E
from minotaur_b200.instances import make_sparse_milp
inst = make_sparse_milp(100_000, 100_000, 10, seed=12345)
eng = E.GpuBoundEngine(0); eng.load_linear(inst)
n = inst.n
h_lb = torch.empty(n, dtype=torch.float64).pin_memory(); h_ub = torch.empty(n, dtype=torch.float64).pin_memory()
opts = E.GpuOptions(E.ROUND_DIRECTED, E.ORDER_JACOBI, E.LOOP_FIXPOINT, 0, E.HANDLERS_ALL)
v = np.zeros(1, np.int32); r = np.zeros(1, np.int32); z = np.zeros(1, np.int64)
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
rl, ru = torch.from_numpy(inst.lb), torch.from_numpy(inst.ub)
tot = []; parts = []
for it in range(25):
    h_lb.copy_(rl); h_ub.copy_(ru); flush.zero_(); torch.cuda.synchronize()
    t = time.perf_counter()
    eng.tighten_raw(1, h_lb.data_ptr(), h_ub.data_ptr(), opts, v.ctypes.data, r.ctypes.data, z.ctypes.data)
    tot.append(time.perf_counter() - t)
    st = eng.stats(); parts.append((st.h2d_ms, st.kernel_ms, st.d2h_ms))
tot = np.array(tot[5:]) * 1e6; parts = np.array(parts[5:]) * 1e3
print(f"e2e wall {tot.mean():.1f} us (min {tot.min():.1f}); device spans: h2d {parts[:,0].mean():.1f}  kernel {parts[:,1].mean():.1f}  d2h {parts[:,2].mean():.1f} us; "
      f"changed vars {int(np.sum(h_lb.numpy() != inst.lb) + 0)} lb / {int(np.sum(h_ub.numpy() != inst.ub))} ub")
