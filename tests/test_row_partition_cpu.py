"""Row-partitioned mode, host side, on CPU: the partitioner, and the per-round merge protocol exercised with
a world_size-2 gloo process group (the oracle's split Jacobi round stands in for the rows / vars kernels).
What it checks is what the NCCL path relies on: merging per-block candidates with an element-wise MAX (lb)
and MIN (ub) all-reduce, then rounding replicated, gives bit for bit the single-process result."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

from minotaur_b200.distributed import partition_rows, row_partition_bounds, shard_boxes
from minotaur_b200.instances import make_sparse_milp


def test_partition_covers_rows_and_balances_nnz():
    inst = make_sparse_milp(1000, 800, 7, seed=3)
    for world in (1, 2, 3, 8):
        cuts = row_partition_bounds(inst.row_ptr, world)
        assert cuts[0] == 0 and cuts[-1] == inst.m and np.all(np.diff(cuts) >= 0)
        blocks = partition_rows(inst, world)
        assert sum(b.m for b in blocks) == inst.m and sum(b.nnz for b in blocks) == inst.nnz
        assert max(b.nnz for b in blocks) <= inst.nnz / world + 7 + 1
        rebuilt = np.concatenate([b.val for b in blocks])
        assert np.array_equal(rebuilt, inst.val)
        for b in blocks:
            b.validate()
    assert shard_boxes(10, 4, 0) == (0, 3) and shard_boxes(10, 4, 3) == (9, 10) and shard_boxes(2, 4, 3) == (2, 2)


# (seed, real data, capacity of a sparse-exchange message: 0 = always the dense all-reduce, 6 = too small in most rounds)
CASES = ((1, False, 0), (2, True, 0), (1, False, 4096), (2, True, 4096), (2, True, 6))


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    from oracle.pyoracle import Oracle
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    orc = Oracle()
    out = []
    for seed, real, cap in CASES:
        inst = make_sparse_milp(400, 350, 6, seed=seed, real_data=real, inf_frac=(0.05, 0.05, 0.0))
        block = partition_rows(inst, world)[rank]
        lb, ub = inst.lb.copy(), inst.ub.copy()
        rounds, verdict, last_changed, sparse_rounds = 0, 0, 0, 0
        while True:
            rounds += 1
            nl, nu, inf = orc.lin_jacobi_round_rows(block, lb, ub)
            merged = False
            if cap > 0:
                # the sparse exchange (mntr_gpu.cu run_rounds_dev): the counts of changed candidates are all-gathered
                # first; if the longest message fits, messages {count, flag | (j, lb, ub)...} of the smallest
                # power-of-two capacity that holds it are all-gathered and merged with max / min; else the dense merge
                idx = np.flatnonzero((nl != lb) | (nu != ub))
                counts = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
                dist.all_gather(counts, torch.tensor([len(idx)], dtype=torch.int64))
                maxc = max(int(c.item()) for c in counts)
                if maxc <= cap:
                    c2 = 4
                    while c2 < maxc:
                        c2 *= 2
                    c2 = min(c2, cap)
                    msg = np.zeros((c2 + 1, 3))
                    msg[0] = (len(idx), float(inf), 0.0)
                    msg[1:len(idx) + 1, 0] = idx; msg[1:len(idx) + 1, 1] = nl[idx]; msg[1:len(idx) + 1, 2] = nu[idx]
                    got = [torch.zeros(c2 + 1, 3, dtype=torch.float64) for _ in range(world)]
                    dist.all_gather(got, torch.from_numpy(msg))
                    ml, mu, flag = nl.copy(), nu.copy(), float(inf)
                    for g in got:
                        c = int(g[0, 0]); flag = max(flag, float(g[0, 1]))
                        j = g[1:c + 1, 0].numpy().astype(np.int64)
                        np.maximum.at(ml, j, g[1:c + 1, 1].numpy()); np.minimum.at(mu, j, g[1:c + 1, 2].numpy())
                    tl = torch.from_numpy(np.concatenate([ml, [flag]])); tu = torch.from_numpy(mu)
                    merged = True
                    sparse_rounds += 1
            if not merged:
                # the dense merge: MAX on lower candidates (+ the row-infeasible flag in an extra slot), MIN on upper
                tl = torch.from_numpy(np.concatenate([nl, [float(inf)]]))
                tu = torch.from_numpy(nu.copy())
                dist.all_reduce(tl, op=dist.ReduceOp.MAX)
                dist.all_reduce(tu, op=dist.ReduceOp.MIN)
            if tl[-1].item() > 0:
                verdict = 2
                break
            new_lb, new_ub, bad, changed = orc.lin_jacobi_round_vars(inst, lb, ub, tl[:-1].numpy(), tu.numpy())
            last_changed = int(np.count_nonzero((new_lb != lb) | (new_ub != ub)))
            lb, ub = new_lb, new_ub
            if bad:
                verdict = 1
                break
            if not changed:
                break
        out.append((lb, ub, verdict, rounds, sparse_rounds))
    q.put((rank, out))
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_merge_protocol_world2_gloo(oracle):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 500)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = dict(q.get(timeout=240) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for k, (seed, real, cap) in enumerate(CASES):
        inst = make_sparse_milp(400, 350, 6, seed=seed, real_data=real, inf_frac=(0.05, 0.05, 0.0))
        jl, ju, jr = oracle.lin_fixpoint_jacobi(inst, inst.lb, inst.ub)
        for rank in (0, 1):
            lb, ub, verdict, rounds, sparse_rounds = results[rank][k]
            assert (sparse_rounds > 0) == (cap == 4096) or cap == 6
            assert (verdict != 0) == (jr["verdict"] != 0)
            if jr["verdict"] == 0:
                assert np.array_equal(lb, jl) and np.array_equal(ub, ju), (seed, rank)
                assert rounds == jr["rounds"]
        # both ranks hold bit-identical boxes
        assert np.array_equal(results[0][k][0], results[1][k][0]) and np.array_equal(results[0][k][1], results[1][k][1])
