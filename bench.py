#!/usr/bin/env python
"""bench.py -- FBBT throughput of the B200 engine on BASELINE.json's configurations.

Headline workload (N=1): config C2 -- synthetic sparse MILP 100k rows x 100k cols, 1M nnz, a single
box tightened to the fixpoint.  A "step" is one whole fixpoint call (all rounds) on the root box.

  value  : nnz-updates/s with the box already resident in HBM (mntr_gpu_tighten_single_dev; CUDA events
           on the engine's stream around the kernel; L2 flushed between steps)
  e2e    : the same metric through the reference-facing C-ABI call mntr_gpu_tighten with PINNED HOST
           buffers: host->device copy of the box, kernel, device->host copy of the result, per step
  roofline: algorithmic HBM bytes of the fixpoint launch (SURVEY.md 8d formula) / launch duration,
           against the measured copy bandwidth in MEASURED_PEAKS.json
  cpu_baseline: the reference's own LinearHandler sweeps (oracle/_ref, built from /root/reference) or,
           if that library is absent, the plain-C oracle port, timed on this host's cores

N>1 (torchrun): every rank tightens its own copy of the C2 instance (independent replicas, weak scaling, no
data-path collective); value = total nnz-updates / max-over-ranks time.  `extra.node_batch` reports the
C3-shaped node batch (8192 boxes sharded by node across the ranks).

`--impl reference` times the reference's CPU implementation on the same config.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

C2 = dict(m=100_000, n=100_000, nnz_per_row=10, seed=12345)
C3 = dict(m=50_000, n=50_000, nnz_per_row=10, seed=2024)
METRIC = "fbbt_nnz_updates_per_sec"
UNIT = "nnz-updates/s"


def measured_hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 100 ms while running."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_baseline_c2(inst, budget_s=12.0):
    """The reference's CPU FBBT on the same instance/box, on this host (1 core: the path is sequential)."""
    from oracle import pyoracle
    if pyoracle.have_reference():
        ref = pyoracle.Reference(inst)
        secs, nnz, reps = 0.0, 0, 0
        while secs < budget_s and reps < 40:
            s, z, _ = ref.time_boxes(0, inst.lb[None, :], inst.ub[None, :])
            secs += s; nnz += z; reps += 1
        ref.close()
        kind = "reference"
        sample = f"{reps} x status-honouring fixpoint of the reference's LinearHandler sweeps on the C2 root box"
    else:
        orc = pyoracle.Oracle()
        secs, nnz, reps = 0.0, 0, 0
        while secs < budget_s and reps < 200:
            t0 = time.perf_counter()
            _, _, r = orc.lin_fixpoint_inplace(inst, inst.lb, inst.ub)
            secs += time.perf_counter() - t0; nnz += r["nnz_updates"]; reps += 1
        kind = "port"
        sample = f"{reps} x in-place fixpoint of the C oracle on the C2 root box"
    return {"value": nnz / secs, "unit": UNIT, "cores": 1, "kind": kind, "sample": sample,
            "host_cores_available": os.cpu_count()}, secs / reps


def run_reference_arm(args, rank, world):
    if rank != 0:
        return
    from minotaur_b200.instances import make_sparse_milp
    from oracle import pyoracle
    inst = make_sparse_milp(**C2)
    have = pyoracle.have_reference()
    if have:
        ref = pyoracle.Reference(inst)
        def step():
            s, z, _ = ref.time_boxes(0, inst.lb[None, :], inst.ub[None, :])
            return s, z
    else:
        orc = pyoracle.Oracle()
        def step():
            t0 = time.perf_counter()
            _, _, r = orc.lin_fixpoint_inplace(inst, inst.lb, inst.ub)
            return time.perf_counter() - t0, r["nnz_updates"]
    for _ in range(args.warmup):
        step()
    secs, nnz = 0.0, 0
    for _ in range(args.steps):
        s, z = step(); secs += s; nnz += z
    val = nnz / secs
    kind = "reference" if have else "port"
    sample = "one status-honouring fixpoint of the reference's LinearHandler sweeps on the C2 root box per step"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * secs / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "C2: synthetic sparse MILP 100k x 100k, 1M nnz, single box to fixpoint", "seed": C2["seed"]},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": 1, "kind": kind, "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-extra", action="store_true", help="skip the C3 node-batch extra measurement")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--batch-boxes", type=int, default=8192)
    ap.add_argument("--profile", action="store_true", help="short run for ncu: no sustained pre-load loop")
    ap.add_argument("--c4-rows-per-rank", type=int, default=2_500_000)
    ap.add_argument("--c5-cons", type=int, default=200_000)
    ap.add_argument("--c5-boxes", type=int, default=1024)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from minotaur_b200 import engine as E
    from minotaur_b200.instances import branch_boxes, make_knapsack_setcover, make_sparse_milp

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the engine has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    # ---------------- workload C2 (one replica per rank) ----------------
    # every rank tightens its own copy of the SAME instance (replicas): per-GPU work is identical, so the job's value
    # over N measures the machine, not the spread of round counts over differently seeded instances
    cfg = dict(C2)
    inst = make_sparse_milp(**cfg)
    eng = E.GpuBoundEngine(local_rank)
    eng.load_linear(inst)
    n = inst.n
    stream = torch.cuda.ExternalStream(eng.stream_handle(), device=dev)
    root_lb = torch.from_numpy(inst.lb).to(dev); root_ub = torch.from_numpy(inst.ub).to(dev)
    w_lb = torch.empty_like(root_lb); w_ub = torch.empty_like(root_ub)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)      # > 126 MB L2

    def dev_step():
        with torch.cuda.stream(stream):
            w_lb.copy_(root_lb); w_ub.copy_(root_ub)
            if not os.environ.get("MNTR_BENCH_NOFLUSH"):
                flush.zero_()                                          # evict L2 between steps
        v, r, z = eng.tighten_single_dev(w_lb.data_ptr(), w_ub.data_ptr())   # synchronises the stream
        st = eng.stats()
        return st.kernel_ms, z, r, v, st

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    # sustained pre-load so the clock samples are taken under load, then the warm-up steps
    t_end = time.perf_counter() + (0.0 if args.profile else 1.5)
    while time.perf_counter() < t_end:
        dev_step()
    for _ in range(args.warmup):
        dev_step()
    barrier()
    t0 = time.perf_counter()
    ker_ms, nnz_tot, st_last = 0.0, 0, None
    for _ in range(args.steps):
        ms, z, r, v, st = dev_step()
        ker_ms += ms; nnz_tot += z; st_last = st
        assert v == 0, "C2 root box must be feasible"
    barrier()
    wall_s = time.perf_counter() - t0
    ker_ms_max = max_over_ranks(ker_ms)
    nnz_all = sum_over_ranks(float(nnz_tot))
    value = nnz_all / (ker_ms_max * 1e-3)

    # ---------------- e2e: C-ABI call with pinned host buffers ----------------
    h_lb = torch.empty(n, dtype=torch.float64).pin_memory(); h_ub = torch.empty(n, dtype=torch.float64).pin_memory()
    opts = E.GpuOptions(E.ROUND_DIRECTED, E.ORDER_JACOBI, E.LOOP_FIXPOINT, 0, E.HANDLERS_ALL)
    vbuf = np.zeros(1, np.int32); rbuf = np.zeros(1, np.int32); zbuf = np.zeros(1, np.int64)
    root_l_h, root_u_h = torch.from_numpy(inst.lb), torch.from_numpy(inst.ub)

    def e2e_step():
        h_lb.copy_(root_l_h); h_ub.copy_(root_u_h)                      # host-side refill, not timed
        with torch.cuda.stream(stream):
            flush.zero_()
        torch.cuda.synchronize()
        t = time.perf_counter()
        eng.tighten_raw(1, h_lb.data_ptr(), h_ub.data_ptr(), opts, vbuf.ctypes.data, rbuf.ctypes.data, zbuf.ctypes.data)
        return time.perf_counter() - t, int(zbuf[0])

    for _ in range(args.warmup):
        e2e_step()
    barrier()
    e2e_s, e2e_nnz = 0.0, 0
    for _ in range(args.steps):
        s, z = e2e_step(); e2e_s += s; e2e_nnz += z
    barrier()
    e2e_value = sum_over_ranks(float(e2e_nnz)) / max_over_ranks(e2e_s)
    # bytes that cross PCIe per step: the kernel reads the whole pinned host box (zero-copy, one coalesced pass) and
    # writes back {lb, ub} of the variables that moved, plus the 128-byte control block
    moved = int(np.count_nonzero((h_lb.numpy() != inst.lb) | (h_ub.numpy() != inst.ub)))
    zero_copy = not os.environ.get("MNTR_GPU_NO_ZEROCOPY")
    d2h_bytes = 16 * moved + 128 if zero_copy else 16 * n + 128
    clocks = sampler.stop() if rank == 0 else None

    # ---------------- roofline of the fixpoint launch ----------------
    peak, peak_src = measured_hbm_peak()
    st = st_last
    algo_bytes = 28 * st.nnz_updates + 24 * st.rows_evaluated + 17 * n * st.max_rounds + 16 * st.n_changes
    launch_ms = ker_ms / args.steps
    achieved = algo_bytes / (launch_ms * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        try:
            traffic = json.load(open(tpath)).get("fbbt_single_jacobi_kernel_dram_bytes_per_launch")
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "kernel": "fbbt_single_jacobi_kernel", "algorithmic_bytes_per_launch": algo_bytes,
                "launch_ms": launch_ms, "rounds": st.max_rounds, "peak_source": peak_src}

    # ---------------- extra: C3-shaped node batch sharded by node ----------------
    extra = {}
    if not args.no_extra:
        for key, fn in (("node_batch", node_batch_extra), ("large_single_box", large_single_extra),
                        ("row_partition", row_partition_extra), ("minlp_batch", minlp_batch_extra)):
            try:
                extra[key] = fn(args, eng, E, torch, dev, stream, rank, world, barrier, max_over_ranks, sum_over_ranks)
            except Exception as ex:  # an extra must never take the headline down
                extra[key] = {"error": repr(ex)[:300]}
            barrier()

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu, _ = cpu_baseline_c2(inst)

    if rank == 0:
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ker_ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": "C2: synthetic sparse MILP 100k x 100k, 1M nnz, single box to fixpoint",
                       "seed": C2["seed"], "rounding": "directed", "order": "jacobi", "l2": "flushed between steps (256 MB write)",
                       "parallelism": f"replicas x{world}" if world > 1 else "single GPU"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 16 * n, "d2h_bytes_per_step": d2h_bytes,
                    "ms_per_step": 1e3 * e2e_s / args.steps, "api": "mntr_gpu_tighten (C ABI), pinned host buffers",
                    "transfer": ("zero-copy: the kernel reads the pinned host box over PCIe and writes back only the "
                                 f"bounds that moved ({moved} variables)") if zero_copy else "staged: cudaMemcpyAsync both ways"},
            "gpu_launches": args.steps, "roofline": roofline, "clocks": clocks,
            "wall_s_timed_region": wall_s, "extra": extra,
        }
        if cpu is not None:
            out["cpu_baseline"] = cpu
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def node_batch_extra(args, eng_c2, E, torch, dev, stream_unused, rank, world, barrier, max_over_ranks, sum_over_ranks):
    """C3: 8192 branching-perturbed boxes on a 50k-row knapsack/set-cover instance, sharded by node."""
    from minotaur_b200.instances import branch_boxes, make_knapsack_setcover
    inst = make_knapsack_setcover(**C3)
    total = args.batch_boxes
    per = (total + world - 1) // world
    b0, b1 = rank * per, min(total, (rank + 1) * per)
    nb = b1 - b0
    # boxes are generated per rank (same seed stream, rank's slice)
    lbs, ubs = branch_boxes(inst.lb, inst.ub, inst.var_type, total, seed=C3["seed"], max_depth=20)
    lbs, ubs = lbs[b0:b1], ubs[b0:b1]
    eng = E.GpuBoundEngine(dev.index)
    eng.load_linear(inst)
    ld = eng.box_ld(nb)
    boxes = torch.empty((inst.n, ld, 2), dtype=torch.float64, device=dev)
    verdict = torch.zeros(ld, dtype=torch.int32, device=dev); rounds = torch.zeros(ld, dtype=torch.int32, device=dev)
    nnz = torch.zeros(ld, dtype=torch.int64, device=dev)
    eng.boxes_upload(lbs, ubs, boxes.data_ptr())
    pristine = boxes.clone()
    ms_tot, reps = 0.0, 3
    for it in range(reps + 1):
        boxes.copy_(pristine)
        torch.cuda.synchronize()
        barrier()
        st = eng.tighten_dev(nb, boxes.data_ptr(), verdict.data_ptr(), rounds.data_ptr(), nnz.data_ptr())
        if it > 0:
            ms_tot += st.kernel_ms
    ms = max_over_ranks(ms_tot / reps)
    nnz_sum = sum_over_ranks(float(nnz[:nb].sum().item()))
    n_inf = sum_over_ranks(float((verdict[:nb] != 0).sum().item()))
    eng.close()
    return {"workload": f"C3: {total} boxes on 50k-row knapsack/set-cover, reference-order sweeps to fixpoint",
            "boxes_per_s": total / (ms * 1e-3), "nnz_updates_per_s": nnz_sum / (ms * 1e-3), "ms": ms,
            "infeasible_boxes": n_inf, "boxes_per_rank": per, "scaling": "strong"}


def large_single_extra(args, eng_c2, E, torch, dev, stream_unused, rank, world, barrier, max_over_ranks,
                       sum_over_ranks):
    """One rank's share of C4 as ONE box on ONE GPU in a single cooperative launch: 2.5M rows x 10 nnz.  Far more rows
    than the grid can keep resident, so K1 streams them from the CSR (dense 32-row blocks lane = row): the
    bandwidth-bound regime of the same kernel whose latency-bound regime is the C2 headline."""
    from minotaur_b200.instances import make_sparse_milp_block
    m = args.c4_rows_per_rank
    inst = make_sparse_milp_block(m, m, 10, seed=777, block=0)
    eng = E.GpuBoundEngine(dev.index)
    eng.load_linear(inst)
    lb0 = torch.from_numpy(inst.lb).to(dev); ub0 = torch.from_numpy(inst.ub).to(dev)
    lb = torch.empty_like(lb0); ub = torch.empty_like(ub0)
    ms, reps, z, r, v = 0.0, 3, 0, 0, 0
    for it in range(reps + 1):
        lb.copy_(lb0); ub.copy_(ub0)
        torch.cuda.synchronize(); barrier()
        v, r, z = eng.tighten_single_dev(lb.data_ptr(), ub.data_ptr())
        st = eng.stats()
        if it > 0:
            ms += st.kernel_ms
    ms = max_over_ranks(ms / reps)
    algo = 28.0 * st.nnz_updates + 24.0 * st.rows_evaluated + 17.0 * m * r + 16.0 * st.n_changes
    peak, _ = measured_hbm_peak()
    eng.close()
    return {"workload": f"{m} rows x {m} cols, {10 * m} nnz, single box to fixpoint in one launch (rows streamed from the CSR)",
            "nnz_updates_per_s": z / (ms * 1e-3), "ms": ms, "rounds": r, "verdict": v,
            "algorithmic_GBps": algo / (ms * 1e-3) / 1e9, "frac_of_hbm_peak": algo / (ms * 1e-3) / 1e9 / peak,
            "scaling": "replicas"}


def row_partition_extra(args, eng_c2, E, torch, dev, stream_unused, rank, world, barrier, max_over_ranks,
                        sum_over_ranks):
    """C4 shape, weak-scaled: every rank owns a block of 2.5M rows x 10 nnz over n = 2.5M x world columns
    (world = 8 is BASELINE config 4: 20M x 20M, 200M nnz), box replicated, per-round NCCL MAX/MIN all-reduce."""
    import torch.distributed as dist
    from minotaur_b200.instances import make_sparse_milp_block
    m_block = args.c4_rows_per_rank
    n = m_block * world
    inst = make_sparse_milp_block(m_block, n, 10, seed=777, block=rank)
    eng = E.GpuBoundEngine(dev.index)
    eng.load_linear(inst)
    if world > 1:
        uid = torch.zeros(128, dtype=torch.uint8, device=dev)
        if rank == 0:
            uid = torch.frombuffer(bytearray(E.GpuBoundEngine.nccl_unique_id()), dtype=torch.uint8).to(dev)
        dist.broadcast(uid, 0)
        eng.comm_init(world, rank, bytes(uid.cpu().numpy().tobytes()))
    lb0 = torch.from_numpy(inst.lb).to(dev); ub0 = torch.from_numpy(inst.ub).to(dev)
    lb = torch.empty_like(lb0); ub = torch.empty_like(ub0)
    best = None
    for it in range(3):
        lb.copy_(lb0); ub.copy_(ub0)
        torch.cuda.synchronize(); barrier()
        t0 = time.perf_counter()
        v, r, z = eng.tighten_single_dev(lb.data_ptr(), ub.data_ptr(),
                                         flags=E.FLAG_PER_ROUND_KERNELS if world == 1 else 0)
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        st = eng.stats()
        rec = dict(wall_ms=1e3 * wall, kernel_ms=st.kernel_ms, rows_ms=st.rows_ms, comm_ms=st.comm_ms,
                   vars_ms=st.vars_ms, rounds=r, verdict=v, nnz=z)
        if it > 0 and (best is None or rec["kernel_ms"] < best["kernel_ms"]):
            best = rec
    ms = max_over_ranks(best["kernel_ms"])
    nnz_total = best["nnz"] if world > 1 else best["nnz"]        # with a communicator nnz is already the job total
    if world > 1:
        eng.comm_destroy()
    eng.close()
    algo_bytes_rank = 28.0 * nnz_total / world + 17.0 * n * best["rounds"]
    return {"workload": f"C4 shape: {m_block * world} rows x {n} cols, {10 * m_block * world} nnz, row-partitioned "
                        f"over {world} GPU(s), single box to fixpoint (Jacobi)",
            "nnz_updates_per_s": nnz_total / (ms * 1e-3), "ms": ms, "rounds": best["rounds"],
            "per_round_ms": {"rows": best["rows_ms"] / best["rounds"], "allreduce": best["comm_ms"] / best["rounds"],
                             "vars": best["vars_ms"] / best["rounds"]},
            "verdict": best["verdict"], "algorithmic_GBps_per_gpu": algo_bytes_rank / (ms * 1e-3) / 1e9,
            "scaling": "weak"}


def minlp_batch_extra(args, eng_c2, E, torch, dev, stream_unused, rank, world, barrier, max_over_ranks,
                      sum_over_ranks):
    """C5 shape (scaled): bilinear/quadratic CGraph constraints + linear rows, node boxes sharded by node."""
    from minotaur_b200.instances import branch_boxes, make_minlp
    n, n_cons, m_lin, total = args.c5_cons, args.c5_cons, args.c5_cons // 10, args.c5_boxes
    lin, tapes = make_minlp(n=n, n_cons=n_cons, m_lin=m_lin, seed=99)
    per = (total + world - 1) // world
    b0, b1 = rank * per, min(total, (rank + 1) * per)
    nb = b1 - b0
    lbs, ubs = branch_boxes(lin.lb, lin.ub, lin.var_type, total, seed=99, max_depth=10, continuous_too=True)
    lbs, ubs = lbs[b0:b1], ubs[b0:b1]
    eng = E.GpuBoundEngine(dev.index)
    eng.load_linear(lin)
    eng.load_cgraph(tapes)
    ld = eng.box_ld(nb)
    boxes = torch.empty((lin.n, ld, 2), dtype=torch.float64, device=dev)
    verdict = torch.zeros(ld, dtype=torch.int32, device=dev); rounds = torch.zeros(ld, dtype=torch.int32, device=dev)
    nnz = torch.zeros(ld, dtype=torch.int64, device=dev)
    eng.boxes_upload(lbs, ubs, boxes.data_ptr())
    pristine = boxes.clone()
    ms_tot, reps = 0.0, 2
    for it in range(reps + 1):
        boxes.copy_(pristine)
        torch.cuda.synchronize(); barrier()
        st = eng.tighten_dev(nb, boxes.data_ptr(), verdict.data_ptr(), rounds.data_ptr(), nnz.data_ptr(),
                             loop=E.LOOP_SIMPLEPRESOLVE)
        if it > 0:
            ms_tot += st.kernel_ms
    ms = max_over_ranks(ms_tot / reps)
    n_inf = sum_over_ranks(float((verdict[:nb] != 0).sum().item()))
    eng.close()
    return {"workload": f"C5 shape (scaled): {n_cons} CGraph constraints + {m_lin} linear rows over {n} variables, "
                        f"{total} node boxes, one presolveNode pass (LinearHandler then NlPresHandler order)",
            "boxes_per_s": total / (ms * 1e-3), "constraint_evals_per_s": 2.0 * total * n_cons / (ms * 1e-3),
            "ms": ms, "infeasible_boxes": n_inf, "boxes_per_rank": per, "scaling": "strong"}


if __name__ == "__main__":
    main()
