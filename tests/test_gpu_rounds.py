"""The per-round kernels of the row-partitioned mode (K5): on one GPU without a communicator they must give
what the single-launch kernel gives (same rounds, same nnz-updates, bounds within 1e-9: the two kernels add a
row's terms in different orders); with 2+ GPUs (threads, one per device, NCCL) every rank must end with the
bit-identical box of the 1-GPU run of the same kernels."""
import threading

import numpy as np
import pytest

from helpers import assert_box_parity
from minotaur_b200 import engine as E
from minotaur_b200.distributed import partition_rows
from minotaur_b200.instances import branch_boxes, make_sparse_milp

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(600)]


@pytest.mark.parametrize("m,n,k,seed,real,inf", [(3000, 2500, 8, 40, False, (0, 0, 0)), (2000, 2200, 12, 41, True, (0.05, 0.05, 0.01)),
                                                 (500, 4000, 40, 42, True, (0, 0, 0))])
def test_per_round_kernels_equal_single_launch(engine, oracle, m, n, k, seed, real, inf):
    inst = make_sparse_milp(m, n, k, seed=seed, real_data=real, inf_frac=inf)
    engine.load_linear(inst)
    lbs, ubs = branch_boxes(inst.lb, inst.ub, inst.var_type, 8, seed=seed, max_depth=12)
    lbs[0], ubs[0] = inst.lb, inst.ub
    for b in range(lbs.shape[0]):
        for rounding in (E.ROUND_NEAREST, E.ROUND_DIRECTED):
            one = engine.tighten(lbs[b], ubs[b], rounding=rounding, order=E.ORDER_JACOBI)
            per = engine.tighten(lbs[b], ubs[b], rounding=rounding, order=E.ORDER_JACOBI, flags=E.FLAG_PER_ROUND_KERNELS)
            assert one.verdict[0] == per.verdict[0]
            if one.verdict[0] == 0:
                assert_box_parity(inst.var_type, per.lb, per.ub, one.lb, one.ub, what=f"box {b} rounding {rounding}")
                assert one.rounds[0] == per.rounds[0] and one.nnz_updates[0] == per.nnz_updates[0]
        jl, ju, jr = oracle.lin_fixpoint_jacobi(inst, lbs[b], ubs[b])
        per = engine.tighten(lbs[b], ubs[b], rounding=E.ROUND_NEAREST, order=E.ORDER_JACOBI, flags=E.FLAG_PER_ROUND_KERNELS)
        assert (per.verdict[0] != 0) == (jr["verdict"] != 0)
        if jr["verdict"] == 0:
            assert_box_parity(inst.var_type, per.lb, per.ub, jl, ju)


def _n_devices():
    return E.load_library().mntr_gpu_device_count()


@pytest.mark.parametrize("xchg", [None, "nccl", "0", "48"], ids=["peer-memory", "nccl-sparse", "dense", "tiny-cap"])
@pytest.mark.parametrize("world", [2, 4, 8])
def test_row_partition_nccl_bitwise_independent_of_ranks(engine, world, xchg, monkeypatch):
    _nccl_partition_case(engine, world, xchg, monkeypatch, ragged=False)


@pytest.mark.parametrize("world", [2, 8])
def test_row_partition_nccl_ragged_rows(engine, world, monkeypatch):
    """The same with rows of 0..40 entries mixed in every block (short rows, tails beyond the 12 staged entries, rows
    longer than a lane takes): whichever rank owns a row, and whatever shares its block, its candidates are the same."""
    _nccl_partition_case(engine, world, None, monkeypatch, ragged=True)


def _nccl_partition_case(engine, world, xchg, monkeypatch, ragged):
    """Row-partitioned mode.  The per-round merge of the candidate bounds is, by default, the exchange over NVLink peer
    memory (every rank pushes its touched candidates into the peers' inboxes; the loop runs on the device); with
    MNTR_GPU_P2P=0 the NCCL sparse exchange (counts all-gathered and read by the host, then right-sized messages
    all-gathered and merged with max / min) or, when the messages would not fit or not pay, the dense MAX/MIN
    all-reduce: MNTR_GPU_SPARSE_XCHG "0" forces the dense merge, a capacity of 48 entries makes most rounds fall back
    to it.  Same bits in every case."""
    if _n_devices() < world:
        pytest.skip(f"needs {world} GPUs")
    monkeypatch.delenv("MNTR_GPU_SPARSE_XCHG", raising=False)
    monkeypatch.delenv("MNTR_GPU_P2P", raising=False)
    if xchg == "nccl":
        monkeypatch.setenv("MNTR_GPU_P2P", "0")
    elif xchg is not None:
        monkeypatch.setenv("MNTR_GPU_SPARSE_XCHG", xchg)      # read by mntr_gpu_comm_init
    if ragged:
        from helpers import ragged_rows
        inst = ragged_rows(make_sparse_milp(20_000, 15_000, 40, seed=78, real_data=True), 78)
    else:
        inst = make_sparse_milp(20_000, 15_000, 9, seed=77, real_data=True, inf_frac=(0.02, 0.02, 0.0))
    engine.load_linear(inst)
    lbs, ubs = branch_boxes(inst.lb, inst.ub, inst.var_type, 3, seed=5, max_depth=10)
    lbs[0], ubs[0] = inst.lb, inst.ub
    ref = [engine.tighten(lbs[b], ubs[b], order=E.ORDER_JACOBI, flags=E.FLAG_PER_ROUND_KERNELS) for b in range(3)]
    blocks = partition_rows(inst, world)
    uid = E.GpuBoundEngine.nccl_unique_id()
    out, errs, sparse = [None] * world, [], [0] * world

    def run(rank):
        try:
            eng = E.GpuBoundEngine(rank)
            eng.load_linear(blocks[rank])
            eng.comm_init(world, rank, uid)
            out[rank] = []
            for b in range(3):
                out[rank].append(eng.tighten(lbs[b], ubs[b], order=E.ORDER_JACOBI))
                sparse[rank] += eng.stats().sparse_rounds
            eng.comm_destroy()
            eng.close()
        except Exception as ex:   # surface worker failures in the main thread
            errs.append((rank, repr(ex)))

    th = [threading.Thread(target=run, args=(r,)) for r in range(world)]
    for t in th:
        t.start()
    for t in th:
        t.join(timeout=300)
    assert not errs, errs
    for b in range(3):
        for rank in range(world):
            got = out[rank][b]
            assert got.verdict[0] == ref[b].verdict[0]
            if ref[b].verdict[0] == 0:
                assert np.array_equal(got.lb, ref[b].lb) and np.array_equal(got.ub, ref[b].ub), (b, rank)
                assert got.rounds[0] == ref[b].rounds[0]
                assert got.nnz_updates[0] == ref[b].nnz_updates[0]     # summed over the ranks
    if xchg in (None, "nccl"):
        assert min(sparse) > 0, "the sparse exchange never ran: the test is vacuous"
    if xchg == "0":
        assert max(sparse) == 0
