"""Sort-based ingest against a literal dict-of-dicts restatement of the reference's build
(dataset.py:199-217,247-271): same id orders, same CSR, same relevant-item CSR — including duplicate (user, item) rows
(first position, last rating), test rows of unknown users (dropped) and test-only items (-1) — and linear-time behaviour."""
import time
from types import SimpleNamespace

import numpy as np
import pandas as pd
import scipy.sparse as sp

from elliot_b200.dataset import DataSet, eval_csr_of, sampler_rows_of, train_csr_of


def _cfg(thr=0):
    return SimpleNamespace(config_test=True, top_k=10, evaluation=SimpleNamespace(simple_metrics=["nDCG"], relevance_threshold=thr))


def _naive(tr, te, thr):
    """The reference's way: {user: {item: rating}} by first appearance, set-comprehension item order."""
    train = {}
    for u, i, r in zip(tr.userId.tolist(), tr.itemId.tolist(), tr.rating.tolist()):
        train.setdefault(u, {})[i] = r
    users = list(train.keys())
    items = list({k for a in train.values() for k in a.keys()})
    pub_u = {u: k for k, u in enumerate(users)}; pub_i = {i: k for k, i in enumerate(items)}
    i_train = {pub_u[u]: {pub_i[i]: r for i, r in its.items()} for u, its in train.items()}
    rows, cols = zip(*[(u, i) for u, its in i_train.items() for i in its])
    csr = sp.csr_matrix((np.ones(len(rows), np.float32), (rows, cols)), shape=(len(users), len(items)))
    test = {}
    for u, i, r in zip(te.userId.tolist(), te.itemId.tolist(), te.rating.tolist()):
        test.setdefault(u, {})[i] = r
    indptr, idx, gain = [0], [], []
    for u in users:
        for it, sc in test.get(u, {}).items():
            if sc >= thr:
                idx.append(pub_i.get(it, -1)); gain.append(2 ** (sc - thr + 1) - 1)
        indptr.append(len(idx))
    ui = [list(set(i_train[u])) for u in range(len(users))]
    return users, items, i_train, csr, (np.array(indptr), np.array(idx, np.int64), np.array(gain, np.float64)), ui, train, test


def _frames(seed, n_users=300, n_items=200, n=6000, dup=True):
    g = np.random.default_rng(seed)
    u = 10 + 3 * g.integers(0, n_users, n); i = 100 + 7 * g.integers(0, n_items, n); r = g.integers(1, 6, n).astype(float)
    tr = pd.DataFrame({"userId": u, "itemId": i, "rating": r})
    if not dup:
        tr = tr.drop_duplicates(["userId", "itemId"]).reset_index(drop=True)
    m = n // 4
    tu = 10 + 3 * g.integers(0, n_users + 20, m)                     # some users unknown to the train split
    ti = 100 + 7 * g.integers(0, n_items + 30, m)                    # some items unknown to the train split
    te = pd.DataFrame({"userId": tu, "itemId": ti, "rating": g.integers(1, 6, m).astype(float)})
    return tr, te


def test_vectorised_build_equals_dict_build():
    for seed, dup, thr in ((0, True, 0), (1, False, 0), (2, True, 3)):
        tr, te = _frames(seed, dup=dup)
        users, items, i_train, csr, ecsr, ui, train, test = _naive(tr, te, thr)
        d = DataSet(_cfg(thr), (tr, te))
        assert d.users == users and d.items == items and d.transactions == csr.nnz
        assert (d.sp_i_train != csr).nnz == 0
        assert d.i_train_dict == i_train and d.train_dict == train
        assert {u: v for u, v in d.test_dict.items()} == {u: test.get(u, {}) for u in users}
        got = eval_csr_of(d, "test")
        for a, b in zip(got, ecsr):
            assert np.array_equal(a, b)
        assert sampler_rows_of(d) == ui
        indptr, flat, srt = train_csr_of(d, "cpu")
        assert np.array_equal(flat.numpy(), np.array([x for r in ui for x in r], np.int32))
        assert np.array_equal(srt.numpy(), np.array([x for r in ui for x in sorted(r)], np.int32))
        # ratings matrix: last rating of a duplicated pair
        for pu in (0, len(users) // 2, len(users) - 1):
            for it, r in i_train[pu].items():
                assert d.sp_i_train_ratings[pu, it] == np.float32(r)


def test_validation_split_and_missing_attribute():
    tr, te = _frames(5)
    d3 = DataSet(_cfg(), (tr, te.iloc[:500], te.iloc[500:]))
    assert eval_csr_of(d3, "val") is not None and set(d3.val_dict.keys()) == set(d3.users)
    d2 = DataSet(_cfg(), (tr, te))
    assert eval_csr_of(d2, "val") is None and not hasattr(d2, "val_dict") and d2.get_validation() is None


def test_ingest_is_fast_at_scale():
    """2 M ratings, 100 K users x 50 K items: ids, CSRs and the device-ready arrays in seconds, no dicts built."""
    g = np.random.default_rng(9)
    n = 2_000_000
    tr = pd.DataFrame({"userId": g.integers(0, 100_000, n), "itemId": g.integers(0, 50_000, n), "rating": np.ones(n)})
    te = pd.DataFrame({"userId": g.integers(0, 100_000, n // 5), "itemId": g.integers(0, 50_000, n // 5), "rating": np.ones(n // 5)})
    t0 = time.perf_counter()
    d = DataSet(_cfg(), (tr, te))
    indptr, flat, srt = train_csr_of(d, "cpu", set_order=False)
    e = eval_csr_of(d, "test")
    dt = time.perf_counter() - t0
    assert flat is None and srt.numel() == d.transactions == d.sp_i_train.nnz and e[0][-1] == e[1].size
    assert not d._lazy, "the throughput path must not materialise the dict views"
    assert dt < 20.0, dt
