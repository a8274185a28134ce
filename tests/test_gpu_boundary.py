"""GPU tests of the boundary additions of round 2: node batches built on the device from deltas, page-locked host
buffers (zero-copy single-box call), the multi-GPU group (one process, several devices), the mod extraction without a
copy of the initial boxes, and the round-tagged control words of the single-launch fixpoint kernel."""
import os
import subprocess
import sys

import numpy as np
import pytest

try:            # before anything loads an NCCL: PyTorch must find ITS bundled libnccl.so.2 first (the engine then reuses it)
    import torch  # noqa: F401
except Exception:   # pragma: no cover
    torch = None

from minotaur_b200 import engine as E
from minotaur_b200.instances import (branch_deltas, deltas_box, make_knapsack_setcover, make_minlp_large, make_sparse_milp,
                                     slice_deltas)

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(600)]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _apply(lb, ub, mv, mu, mx):
    lb, ub = lb.copy(), ub.copy()
    u = mu.astype(bool)
    ub[mv[u]] = mx[u]
    lb[mv[~u]] = mx[~u]
    return lb, ub


def test_boxes_from_deltas_matches_dense_upload(engine):
    inst = make_knapsack_setcover(3000, 2500, 8, seed=5)
    engine.load_linear(inst)
    nb = 77
    d = branch_deltas(inst.lb, inst.ub, inst.var_type, nb, seed=3, max_depth=12)
    ld = engine.box_ld(nb)
    a = torch.zeros((inst.n, ld, 2), dtype=torch.float64, device="cuda")
    engine.boxes_from_deltas(inst.lb, inst.ub, *d, a.data_ptr())
    lb, ub = engine.boxes_download(nb, a.data_ptr())
    for b in range(nb):
        l, u = deltas_box(inst.lb, inst.ub, d, b)
        assert np.array_equal(lb[b], l) and np.array_equal(ub[b], u), b


@pytest.mark.parametrize("loosen", [False, True])
def test_tighten_nodes_mods_without_initial_copy(engine, oracle, loosen):
    """The mods are the final bounds that differ from root + deltas; the kernels compare against the ROOT box and search
    the box's delta list only where needed.  `loosen`: some deltas LOOSEN the root bound, so a final bound may equal the
    root's and still be a mod; repeated deltas on one (variable, side): the last one defines the initial bound."""
    inst = make_sparse_milp(700, 600, 6, seed=70, real_data=True)
    engine.load_linear(inst)
    nb = 64
    ptr, var, up, val = branch_deltas(inst.lb, inst.ub, inst.var_type, nb, seed=8, max_depth=10, continuous_too=True)
    if loosen:
        rng = np.random.default_rng(1)
        val = val.copy()
        pick = rng.random(len(val)) < 0.3
        val[pick] = np.where(up[pick] == 1, inst.ub[var[pick]] + 2.0, inst.lb[var[pick]] - 2.0)     # looser than the root
        # duplicate the first delta of every non-empty box with another value in front of it
        keep = np.nonzero(np.diff(ptr) > 0)[0]
        ins = ptr[keep]
        var = np.insert(var, ins, var[ins]); up = np.insert(up, ins, up[ins]); val = np.insert(val, ins, val[ins] + 0.25)
        add = np.zeros(nb + 1, np.int64); add[keep + 1] = 1
        ptr = ptr + np.cumsum(add)
    d = (ptr, var, up, val)
    v, r, mp, mv, mu, mx, total = engine.tighten_nodes(inst.lb, inst.ub, *d, rounding=E.ROUND_NEAREST)
    assert total == mp[-1] == len(mv)
    n_feas = 0
    for b in range(nb):
        lb0, ub0 = deltas_box(inst.lb, inst.ub, d, b)
        ol, ou, ro = oracle.lin_fixpoint_inplace(inst, lb0, ub0)
        assert (v[b] != 0) == (ro["verdict"] != 0), b
        a, e = int(mp[b]), int(mp[b + 1])
        if v[b] != 0:
            assert a == e
            continue
        n_feas += 1
        keys = list(zip(mv[a:e].tolist(), mu[a:e].tolist()))
        assert keys == sorted(set(keys)), b
        gl, gu = _apply(lb0, ub0, mv[a:e], mu[a:e], mx[a:e])
        assert np.array_equal(gl, ol) and np.array_equal(gu, ou), b
        assert e - a == int(np.count_nonzero(ol != lb0) + np.count_nonzero(ou != ub0)), b
    assert n_feas > nb // 4


@pytest.mark.parametrize("dirty_root", [False, True])
def test_prepared_batch_equals_uploaded_batch(engine, oracle, dirty_root):
    """A batch built on the device from root + deltas is PREPARED (BatchIo::prepared): the first sweep's integer rounding
    and bound check visit only the variables the deltas set and the ones the rows moved.  The result must be the
    uploaded-boxes path's (which runs the all-variables pass) and the oracle's, bit for bit -- with deltas that put
    FRACTIONAL bounds on integer variables, deltas that CROSS a variable's bounds (checkBounds_ must report them), and,
    for `dirty_root`, a root box that itself holds fractional integer bounds (the prepared path must step aside)."""
    inst = make_knapsack_setcover(1200, 1000, 8, seed=11)
    engine.load_linear(inst)
    root_lb, root_ub = inst.lb.copy(), inst.ub.copy()
    isint = np.nonzero(inst.var_type <= 1)[0]
    if dirty_root:
        root_ub[isint[::7]] += 0.5                      # ub of an integer variable = k + 0.5: tightenInts_ floors it
    nb = 70
    ptr, var, up, val = branch_deltas(root_lb, root_ub, inst.var_type, nb, seed=4, max_depth=8)
    rng = np.random.default_rng(12)
    val = val.copy()
    frac = rng.random(len(val)) < 0.25
    val[frac] += np.where(up[frac] == 1, 0.5, -0.5)       # fractional, looser by half a unit: rounding restores the integer
    # two boxes whose deltas cross a variable's bounds outright
    crossed = []
    for b in (5, 40):
        q = int(ptr[b])
        if ptr[b + 1] > q:
            j = var[q]
            val[q] = root_lb[j] - 3.0 if up[q] == 1 else root_ub[j] + 3.0
            if not np.any(var[q + 1:ptr[b + 1]] == j): crossed.append(b)
    d = (ptr, var, up, val)
    v, r, mp, mv, mu, mx, total = engine.tighten_nodes(root_lb, root_ub, *d, rounding=E.ROUND_NEAREST)
    L = np.empty((nb, inst.n)); U = np.empty((nb, inst.n))
    for b in range(nb):
        L[b], U[b] = deltas_box(root_lb, root_ub, d, b)
    up_res = engine.tighten(L, U, rounding=E.ROUND_NEAREST, order=E.ORDER_REFERENCE)
    assert np.array_equal(v != 0, up_res.verdict != 0)
    assert np.array_equal(r, up_res.rounds)
    n_feas = 0
    for b in range(nb):
        ol, ou, ro = oracle.lin_fixpoint_inplace(inst, L[b], U[b])
        assert (v[b] != 0) == (ro["verdict"] != 0), b
        if v[b] != 0:
            continue
        n_feas += 1
        a, e = int(mp[b]), int(mp[b + 1])
        gl, gu = _apply(L[b], U[b], mv[a:e], mu[a:e], mx[a:e])
        assert np.array_equal(gl, up_res.lb[b]) and np.array_equal(gu, up_res.ub[b]), b
        assert np.array_equal(gl, ol) and np.array_equal(gu, ou), b
    assert crossed and all(v[b] != 0 for b in crossed) and n_feas > nb // 4


def test_zero_copy_call_on_alloc_host_buffers(engine, oracle):
    inst = make_sparse_milp(5000, 5000, 8, seed=21)
    engine.load_linear(inst)
    ref = engine.tighten(inst.lb, inst.ub)                       # pageable numpy buffers: staged copies
    h_lb = engine.alloc_host(inst.n); h_ub = engine.alloc_host(inst.n)
    try:
        h_lb[:] = inst.lb; h_ub[:] = inst.ub
        o = E.GpuOptions(E.ROUND_DIRECTED, E.ORDER_JACOBI, E.LOOP_FIXPOINT, 0, E.HANDLERS_ALL)
        v = np.zeros(1, np.int32); r = np.zeros(1, np.int32); z = np.zeros(1, np.int64)
        engine.tighten_raw(1, h_lb.ctypes.data, h_ub.ctypes.data, o, v.ctypes.data, r.ctypes.data, z.ctypes.data)
        assert v[0] == ref.verdict[0] == 0 and r[0] == ref.rounds[0] and z[0] == ref.nnz_updates[0]
        assert np.array_equal(h_lb, ref.lb) and np.array_equal(h_ub, ref.ub)
    finally:
        engine.free_host(h_lb); engine.free_host(h_ub)


def _group_case(devices, oracle):
    lin, tapes = make_minlp_large(3000, 3000, 300, seed=5)
    nb = 200
    d = branch_deltas(lin.lb, lin.ub, lin.var_type, nb, seed=6, max_depth=8, continuous_too=True)
    with E.GpuBoundGroup(devices) as g, E.GpuBoundEngine(devices[0]) as one:
        assert g.size == len(devices)
        g.load_linear(lin); g.load_cgraph(tapes)
        one.load_linear(lin); one.load_cgraph(tapes)
        kw = dict(rounding=E.ROUND_NEAREST, loop=E.LOOP_SIMPLEPRESOLVE)
        a = g.tighten_nodes(lin.lb, lin.ub, *d, **kw)
        b = one.tighten_nodes(lin.lb, lin.ub, *d, **kw)
        for x, y in zip(a[:6], b[:6]):
            assert np.array_equal(x, y)
        assert a[6] == b[6]
        # too small a buffer: the total is still the batch's
        c = g.tighten_nodes(lin.lb, lin.ub, *d, mod_cap=5, **kw)
        assert c[6] == a[6] and np.array_equal(c[2], a[2])
    o = oracle.batch_deltas(lin, tapes, 2, lin.lb, lin.ub, d, n_threads=4, mod_cap=1 << 13)
    v, r, mp, mv, mu, mx, _ = a
    n_cmp = 0
    for k in range(nb):
        if v[k] == E.INFEAS_ROW:
            continue
        assert (v[k] != 0) == (o["verdict"][k] != 0), k
        if v[k] == 0:
            cnt = int(o["mod_cnt"][k]); s, e = int(mp[k]), int(mp[k + 1])
            assert e - s == cnt and np.array_equal(mv[s:e], o["mod_var"][k, :cnt]) and np.array_equal(mx[s:e], o["mod_val"][k, :cnt]), k
            n_cmp += 1
    assert n_cmp > 20


def test_group_of_one_device_equals_single_context(oracle):
    _group_case([0], oracle)


@pytest.mark.parametrize("n_dev", [2, 8])
def test_group_shards_a_node_batch_over_devices(oracle, n_dev):
    lib = E.load_library()
    if lib.mntr_gpu_device_count() < n_dev:
        pytest.skip(f"needs {n_dev} GPUs")
    _group_case(list(range(n_dev)), oracle)


STRESS = r"""
import sys, numpy as np
sys.path.insert(0, %r)
from minotaur_b200 import engine as E
from minotaur_b200.instances import make_sparse_milp, branch_boxes
from oracle.pyoracle import Oracle
orc = Oracle()
inst = make_sparse_milp(6000, 6000, 8, seed=44)
lbs, ubs = branch_boxes(inst.lb, inst.ub, inst.var_type, 6, seed=2, max_depth=12)
lbs[0], ubs[0] = inst.lb, inst.ub
with E.GpuBoundEngine(0) as eng:
    eng.load_linear(inst)
    for flags in (0, E.FLAG_STAGED_ROWS):
        for loop in (E.LOOP_FIXPOINT, E.LOOP_SIMPLEPRESOLVE):
            for b in range(lbs.shape[0]):
                res = eng.tighten(lbs[b], ubs[b], order=E.ORDER_JACOBI, loop=loop, flags=flags)
                if loop == E.LOOP_FIXPOINT:
                    jl, ju, jr = orc.lin_fixpoint_jacobi(inst, lbs[b], ubs[b])
                    assert (res.verdict[0] != 0) == (jr["verdict"] != 0), (flags, b)
                    if jr["verdict"] == 0:
                        assert res.rounds[0] == jr["rounds"] and res.nnz_updates[0] == jr["nnz_updates"], (flags, b)
                        assert np.allclose(res.lb, jl, rtol=1e-9, atol=1e-9) and np.allclose(res.ub, ju, rtol=1e-9, atol=1e-9)
print("stress ok")
"""


def test_barrier_snapshot_is_immune_to_next_round_writers():
    """a block that takes its snapshot of barrier r late may see the control words of round r+1 written by
    faster blocks.  MNTR_GPU_STRESS_BARRIER makes block 1's poller sleep 40 us after arriving at every barrier, so the
    other blocks are deep in the next round (or past it) when it looks: results must not change and nothing may hang.
    Run in a subprocess under a timeout: a hang must fail the test, not the session."""
    env = dict(os.environ, MNTR_GPU_STRESS_BARRIER="40000")
    p = subprocess.run([sys.executable, "-c", STRESS % ROOT], env=env, capture_output=True, text=True, timeout=300)
    assert p.returncode == 0 and "stress ok" in p.stdout, p.stdout[-2000:] + p.stderr[-2000:]
