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
    for seed, real in ((1, False), (2, True)):
        inst = make_sparse_milp(400, 350, 6, seed=seed, real_data=real, inf_frac=(0.05, 0.05, 0.0))
        block = partition_rows(inst, world)[rank]
        lb, ub = inst.lb.copy(), inst.ub.copy()
        rounds, verdict = 0, 0
        while True:
            rounds += 1
            nl, nu, inf = orc.lin_jacobi_round_rows(block, lb, ub)
            # the merge: MAX on lower candidates (+ the row-infeasible flag in an extra slot), MIN on upper
            tl = torch.from_numpy(np.concatenate([nl, [float(inf)]]))
            tu = torch.from_numpy(nu.copy())
            dist.all_reduce(tl, op=dist.ReduceOp.MAX)
            dist.all_reduce(tu, op=dist.ReduceOp.MIN)
            if tl[-1].item() > 0:
                verdict = 2
                break
            lb, ub, bad, changed = orc.lin_jacobi_round_vars(inst, lb, ub, tl[:-1].numpy(), tu.numpy())
            if bad:
                verdict = 1
                break
            if not changed:
                break
        out.append((lb, ub, verdict, rounds))
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
    for k, (seed, real) in enumerate(((1, False), (2, True))):
        inst = make_sparse_milp(400, 350, 6, seed=seed, real_data=real, inf_frac=(0.05, 0.05, 0.0))
        jl, ju, jr = oracle.lin_fixpoint_jacobi(inst, inst.lb, inst.ub)
        for rank in (0, 1):
            lb, ub, verdict, rounds = results[rank][k]
            assert (verdict != 0) == (jr["verdict"] != 0)
            if jr["verdict"] == 0:
                assert np.array_equal(lb, jl) and np.array_equal(ub, ju), (seed, rank)
                assert rounds == jr["rounds"]
        # both ranks hold bit-identical boxes
        assert np.array_equal(results[0][k][0], results[1][k][0]) and np.array_equal(results[0][k][1], results[1][k][1])
