"""Root-presolve row operations: duplicate-row candidates of LinearHandler::dupRows_ and rows that are
redundant on a box (linBndTighten_ root mode).  CPU: the oracle against the reference's own dupRows_ / getLfBnds_ run
through the harness (oracle/_ref; it travels prebuilt to the GPU box).  GPU: the engine against the oracle, bit for
bit (hashes, candidate list, flags)."""
import numpy as np
import pytest

from minotaur_b200.instances import branch_boxes, make_sparse_milp, plant_duplicate_rows


def _case(seed, m=300, n=200, k=5):
    return plant_duplicate_rows(make_sparse_milp(m, n, k, seed=seed, real_data=(seed % 2 == 1)), 40, seed)


@pytest.mark.parametrize("seed", [0, 1, 2, 3])
def test_oracle_dup_rows_reproduce_reference(oracle, have_ref, seed):
    """The oracle's candidate list, walked in order with the reference's own treatDupRows_, leaves the problem exactly as
    LinearHandler::dupRows_ leaves it (same rows deleted, same merged row bounds)."""
    if not have_ref:
        pytest.skip("oracle/_ref not built")
    from oracle.pyoracle import Reference
    inst = _case(seed)
    a, b = Reference(inst), Reference(inst)
    r1, r2 = a.draw_dup_vectors(4321 + seed)
    dA, lA, uA = a.dup_rows(4321 + seed, inst.m)
    h1, h2, pairs = oracle.root_dup_rows(inst, r1, r2)
    dB, lB, uB = b.dup_rows_replay(pairs, h1, inst.m)
    a.close(); b.close()
    assert dA.sum() >= 20
    assert np.array_equal(dA, dB) and np.array_equal(lA, lB) and np.array_equal(uA, uB)


def test_oracle_redundant_rows_match_reference(oracle, have_ref):
    if not have_ref:
        pytest.skip("oracle/_ref not built")
    from oracle.pyoracle import Reference
    inst = _case(7)
    ref = Reference(inst)
    lbs, ubs = branch_boxes(inst.lb, inst.ub, inst.var_type, 6, seed=3, max_depth=40, continuous_too=True)
    n = 0
    for b in range(6):
        o = oracle.root_redundant_rows(inst, lbs[b], ubs[b])
        assert np.array_equal(o, ref.redundant_rows(lbs[b], ubs[b], inst.m)), b
        n += int(o.sum())
    ref.close()
    assert n > 20


@pytest.mark.gpu
@pytest.mark.timeout(600)
@pytest.mark.parametrize("seed,m,n,k", [(0, 300, 200, 5), (1, 5000, 3000, 8), (2, 777, 64, 3)])
def test_gpu_root_rows_vs_oracle(engine, oracle, seed, m, n, k):
    inst = plant_duplicate_rows(make_sparse_milp(m, n, k, seed=seed, real_data=(seed % 2 == 1)), max(40, m // 20), seed)
    engine.load_linear(inst)
    rng = np.random.default_rng(seed)
    r1, r2 = rng.random(n) * 10.0, rng.random(n) * 10.0
    h1, h2, pairs = engine.root_dup_rows(r1, r2, cap=16)            # a small capacity first: the call reports the total
    o1, o2, opairs = oracle.root_dup_rows(inst, r1, r2, cap=1 << 18)
    assert np.array_equal(h1, o1) and np.array_equal(h2, o2)
    assert pairs.shape == opairs.shape and np.array_equal(pairs, opairs)
    assert len(pairs) >= 20
    lbs, ubs = branch_boxes(inst.lb, inst.ub, inst.var_type, 4, seed=seed, max_depth=40, continuous_too=True)
    for b in range(4):
        assert np.array_equal(engine.root_redundant_rows(lbs[b], ubs[b]), oracle.root_redundant_rows(inst, lbs[b], ubs[b])), b


@pytest.mark.gpu
@pytest.mark.timeout(600)
def test_gpu_dup_rows_at_c2_size(engine):
    """100k rows: the all-pairs compare (5e9 pairs) on the device; planted duplicates must all be among the candidates."""
    inst = plant_duplicate_rows(make_sparse_milp(100_000, 100_000, 10, seed=12345), 500, 9)
    engine.load_linear(inst)
    rng = np.random.default_rng(1)
    r1, r2 = rng.random(inst.n) * 10.0, rng.random(inst.n) * 10.0
    h1, h2, pairs = engine.root_dup_rows(r1, r2)
    st = engine.stats()
    assert len(pairs) >= 400
    # every returned pair passes the reference's test on the returned hashes
    i, j = pairs[:, 0], pairs[:, 1]
    with np.errstate(divide="ignore", invalid="ignore"):
        same = (np.abs(h1[j] - h1[i]) < 1e-10) | (np.abs(h1[j] + h1[i]) < 1e-10)
        mult = np.abs(h1[i] / h1[j] - h2[i] / h2[j]) < 1e-10
    assert np.all(i < j) and np.all(same | mult)
    assert np.all((pairs[:, 2] == 1) == same)
    print(f"dupRows_ all-pairs compare of 100k rows: {st.kernel_ms:.2f} ms on the device, {len(pairs)} candidates")


# ---- LinearHandler::coeffImp_ (LinearHandler.cpp:600-704) ----

def _same_improvements(a, b):
    return all(len(x) == len(y) and np.array_equal(x, y) for x, y in zip(a[:5], b[:5]))


@pytest.mark.parametrize("seed", range(6))
def test_oracle_coeff_imp_matches_reference(oracle, have_ref, seed):
    """The restatement against the reference's own coeffImp_ (called through the LinearHandler subclass of the harness) on
    big-M instances: every improved row, the variable, the new coefficient (bitwise), the row bound that moved."""
    from minotaur_b200.instances import make_bigm_instance
    if not have_ref:
        pytest.skip("oracle/_ref not built")
    from oracle import pyoracle
    inst = make_bigm_instance(40, 120, 400, seed)
    o = oracle.root_coeff_imp(inst, inst.lb, inst.ub)
    r = pyoracle.Reference(inst).coeff_imp(inst.lb, inst.ub)
    assert len(o[0]) > 50 and set(np.unique(o[3])) == {0, 1, 2}          # all four cases of :645-694 occur
    assert _same_improvements(o, r)


@pytest.mark.gpu
@pytest.mark.parametrize("seed,sizes", [(0, (40, 120, 400)), (1, (10, 30, 90)), (2, (300, 900, 4000)), (3, (2000, 6000, 30000))])
def test_gpu_coeff_imp_vs_oracle(oracle, seed, sizes):
    """mntr_gpu_root_coeff_imp against the oracle: same rows, variables, coefficients (bitwise) and row bounds; the rows
    that read 2-term rows improved before them run in a later dependency level."""
    from minotaur_b200 import engine as E
    from minotaur_b200.instances import make_bigm_instance
    inst = make_bigm_instance(*sizes, seed)
    eng = E.GpuBoundEngine(0)
    g = eng.root_coeff_imp(inst, inst.lb, inst.ub)
    eng.close()
    o = oracle.root_coeff_imp(inst, inst.lb, inst.ub)
    assert len(o[0]) > 10
    assert g[5]["levels"] >= 2 and g[5]["count"] == len(o[0])
    assert _same_improvements(g, o)
