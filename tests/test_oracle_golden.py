"""The oracle restatement (oracle/bprmf_oracle.c) against goldens minted from the reference's
own code (oracle/gen_golden.py) and against numpy's legacy RandomState."""
import os

import numpy as np
import pytest

import oracle

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_mt19937_raw_stream_matches_numpy():
    rs = np.random.RandomState(42)
    want = rs.randint(0, 2 ** 32, size=2000, dtype=np.uint64).astype(np.uint32)
    got = oracle.Rng(42).raw(2000)
    assert np.array_equal(got, want)


def test_randint_masked_rejection_matches_numpy():
    for n in [1, 2, 3, 132, 3706, 6040, 100000, 1000000, 2 ** 20, 2 ** 31 - 1]:
        np.random.seed(7)
        want = [int(np.random.randint(n)) for _ in range(200)]
        r = oracle.Rng(7)
        got = [r.randint(n) for _ in range(200)]
        assert got == want, n


def test_legacy_normal_matches_numpy():
    np.random.seed(123)
    want = np.random.normal(0.0, 0.1, size=(37, 5))
    got = oracle.Rng(123).normal(0.0, 0.1, (37, 5))
    assert np.array_equal(got, want)


def test_mf_init_bit_exact(golden):
    U, V, b = oracle.mf_init(int(golden["model_seed"]), len(golden["users"]), len(golden["items"]), int(golden["d"]))
    assert np.array_equal(U, golden["U0"]) and np.array_equal(V, golden["V0"]) and not b.any()


def _replay(g):
    nu, ni = len(g["users"]), len(g["items"])
    rng = oracle.Rng(42)
    T, E = int(g["transactions"]), int(g["epochs"])
    out = []
    for _ in range(E):
        out.append(oracle.sampler_step(rng, nu, ni, g["ui_indptr"], g["ui_indices"], T))
    return rng, out


def test_sampler_stream_bit_exact_across_epochs(golden):
    rng, eps = _replay(golden)
    tu = np.concatenate([e[0] for e in eps]); ti = np.concatenate([e[1] for e in eps]); tj = np.concatenate([e[2] for e in eps])
    assert np.array_equal(tu, golden["tu"]) and np.array_equal(ti, golden["ti"]) and np.array_equal(tj, golden["tj"])
    assert [rng.randint(1 << 20) for _ in range(4)] == list(golden["tail"])


def test_sequential_update_matches_reference(golden):
    g = golden
    U, V, b = g["U0"].copy(), g["V0"].copy(), np.zeros(len(g["items"]))
    T = int(g["transactions"])
    oracle.bpr_update_seq(U, V, b, g["tu"][:T], g["ti"][:T], g["tj"][:T], *g["hp"])
    assert np.abs(U - g["U_ep1"]).max() < 1e-13 and np.abs(V - g["V_ep1"]).max() < 1e-13
    assert np.abs(b - g["b_ep1"]).max() < 1e-13
    oracle.bpr_update_seq(U, V, b, g["tu"][T:], g["ti"][T:], g["tj"][T:], *g["hp"])
    assert np.abs(U - g["U"]).max() < 1e-13 and np.abs(V - g["V"]).max() < 1e-13 and np.abs(b - g["b"]).max() < 1e-13


def test_topk_matches_reference(golden):
    g = golden
    idx, val = oracle.user_topk(g["U"], g["V"], g["b"], g["ui_indptr"], g["ui_indices"],
                                np.arange(len(g["users"])), int(g["k"]))
    assert np.array_equal(idx, g["rec_idx"])
    fin = np.isfinite(g["rec_val"])
    assert np.abs(val - g["rec_val"])[fin].max() < 1e-12


def test_sampler_rejects_user_owning_everything():
    import pytest
    indptr = np.array([0, 3], np.int64); indices = np.array([0, 1, 2], np.int32)
    with pytest.raises(RuntimeError):
        oracle.sampler_step(oracle.Rng(42), 1, 3, indptr, indices, 5)


def test_numpy_port_matches_golden(golden_tiny):
    """The interpreter-bound NumPy port used as the reference-speed CPU baseline replays the
    reference's stream and tables exactly."""
    from oracle import bprmf_numpy as bn
    g = golden_tiny
    rows = [list(g["ui_indices"][g["ui_indptr"][u]:g["ui_indptr"][u + 1]]) for u in range(len(g["users"]))]
    m = bn.SequentialBPR(len(g["users"]), len(g["items"]), int(g["d"]), *[float(x) for x in g["hp"]],
                         seed=int(g["model_seed"]))
    assert np.array_equal(m.P, g["U0"])
    np.random.seed(42)
    T = int(g["transactions"])
    got = []
    for u, i, j in bn.triple_stream(rows, len(g["items"]), T):
        got.append((u, i, j)); m.sgd(u, i, j)
    got = np.array(got)
    assert np.array_equal(got[:, 0], g["tu"][:T]) and np.array_equal(got[:, 1], g["ti"][:T]) and np.array_equal(got[:, 2], g["tj"][:T])
    assert np.abs(m.P - g["U_ep1"]).max() < 1e-13 and np.abs(m.Q - g["V_ep1"]).max() < 1e-13


# ---------------------------------------------------------------- MF2020 (pointwise, NumPy reference -> pinned)
@pytest.mark.parametrize("case", ["tiny", "small"])
def test_mf2020_oracle_matches_reference_run(case):
    """oracle.mf2020_epoch_samples / mf2020_update_seq against the reference's own MF2020 run
    (tests/golden/mf2020_<case>.npz minted by oracle/gen_golden.py): init bit-exact, sample lists bit-exact,
    tables / biases / batch losses to 1e-12."""
    import random
    g = dict(np.load(os.path.join(GOLDEN, f"mf2020_{case}.npz")))
    gb_ = dict(np.load(os.path.join(GOLDEN, f"bprmf_{case}.npz")))
    nu, ni, d, seed, m = len(gb_["users"]), len(gb_["items"]), int(g["d"]), int(g["seed"]), int(g["m"])
    rs = np.random.RandomState(seed); pr = random.Random(seed)
    U = rs.normal(0, 0.1, (nu, d)); V = rs.normal(0, 0.1, (ni, d))
    assert np.array_equal(U, g["U0"]) and np.array_equal(V, g["V0"])
    ub, ib, gbias = np.zeros(nu), np.zeros(ni), 0.0
    pos = g["positives"]
    for ep in range(int(g["epochs"])):
        smp = oracle.mf2020_epoch_samples(rs, pr, pos[:, 0], pos[:, 1], ni, m)
        assert np.array_equal(smp, g[f"samples_ep{ep}"])
        gbias, bl = oracle.mf2020_update_seq(U, V, ub, ib, gbias, smp[:, 0], smp[:, 1], smp[:, 2], float(g["lr"]), float(g["reg"]))
        assert np.allclose(bl, g[f"batch_loss_ep{ep}"], rtol=1e-12, atol=0)
    assert np.abs(U - g["U"]).max() < 1e-12 and np.abs(V - g["V"]).max() < 1e-12
    assert np.abs(ub - g["ub"]).max() < 1e-12 and np.abs(ib - g["ib"]).max() < 1e-12 and abs(gbias - float(g["gb"])) < 1e-12
