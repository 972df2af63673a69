"""CPU tests of the host-side mirror against goldens minted from the reference's own DataSet,
Splitter, Sampler and Evaluator (oracle/gen_golden.py)."""
import os
from types import SimpleNamespace

import numpy as np
import pandas as pd
import pytest

from elliot_b200.dataset import DataSet
from elliot_b200.evaluation import Evaluator
from elliot_b200.recommender.early_stopping import EarlyStopping
from elliot_b200.run import split_random_subsampling

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _frames(g):
    f = lambda a: pd.DataFrame({"userId": a[:, 0].astype(np.int64), "itemId": a[:, 1].astype(np.int64), "rating": a[:, 2]})
    return f(g["train"]), f(g["test"])


def _config(k):
    return SimpleNamespace(config_test=False, top_k=k,
                           evaluation=SimpleNamespace(simple_metrics=["nDCG", "HR", "Precision", "Recall"],
                                                      relevance_threshold=0, paired_ttest=False, cutoffs=[k]))


def test_dataset_id_order_and_sampler_rows_match_reference(golden):
    g = golden
    data = DataSet(_config(int(g["k"])), _frames(g))
    assert data.users == list(g["users"]) and data.items == list(g["items"])        # dataset.py:201-202 ordering
    assert data.transactions == int(g["transactions"])
    rows = data.sampler_rows()                                                        # custom_sampler.py:21 ordering
    flat = np.array([x for r in rows for x in r], np.int32)
    assert np.array_equal(flat, g["ui_indices"])
    assert np.array_equal(np.cumsum([0] + [len(r) for r in rows]), g["ui_indptr"])
    assert data.sp_i_train.nnz == data.transactions and data.allunrated_mask.shape == (data.num_users, data.num_items)
    assert (~data.allunrated_mask).sum() == data.transactions


def test_evaluator_matches_reference_metrics(golden):
    """nDCG/HR/Precision/Recall on the reference's own top-k lists == reference Evaluator output."""
    g = golden
    k = int(g["k"])
    data = DataSet(_config(k), _frames(g))
    ev = Evaluator(data, SimpleNamespace(meta=SimpleNamespace()))
    recs = {}
    for pu, u in enumerate(data.users):
        recs[u] = [(data.items[i], float(v)) for i, v in zip(g["rec_idx"][pu], g["rec_val"][pu]) if i >= 0]
    got = ev.eval((recs, recs))[k]["test_results"]
    want = dict(zip(g["metric_names"].tolist(), g["metric_vals"].tolist()))
    for m in want:
        assert abs(got[m] - want[m]) < 1e-12, (m, got[m], want[m])


def test_random_subsampling_split_matches_reference():
    g = np.load(os.path.join(GOLDEN, "split_random_subsampling.npz"))
    a = g["data"]
    df = pd.DataFrame({"userId": a[:, 0].astype(np.int64), "itemId": a[:, 1].astype(np.int64), "rating": a[:, 2]})
    (train, test), = split_random_subsampling(df, 0.2, 42)
    assert np.array_equal(train[["userId", "itemId", "rating"]].to_numpy(), g["train"])
    assert np.array_equal(test[["userId", "itemId", "rating"]].to_numpy(), g["test"])


def test_early_stopping_rules():
    mk = lambda **kw: EarlyStopping(SimpleNamespace(**kw), "nDCG", 10, [10], ["nDCG"])
    res = lambda vals: [{10: {"val_results": {"nDCG": v}}} for v in vals]
    assert not mk().stop([], res([0.1, 0.05]))                       # inactive without options
    es = mk(patience=1)
    assert not es.stop([], res([0.1]))
    assert not es.stop([], res([0.1, 0.2]))                          # still improving
    assert es.stop([], res([0.3, 0.2, 0.1]))                         # two consecutive drops
    assert not es.stop([], res([0.3, 0.1, 0.2]))
    assert not mk(patience=0).stop([], res([0.3, 0.2, 0.1]))         # reference quirk: a 1-value window has no pair
    el = mk(patience=1, monitor="loss")
    assert el.stop([1.0, 2.0], []) and not el.stop([2.0, 1.0], [])
    em = mk(patience=1, min_delta=0.05)
    assert em.stop([], res([0.10, 0.12]))                            # improvement below min_delta counts as stall
    with pytest.raises(Exception):
        mk(monitor="MAP")


def test_base_model_contract_errors():
    from elliot_b200.recommender.base_recommender_model import BaseRecommenderModel

    class M(BaseRecommenderModel):
        train = get_recommendations = get_loss = get_params = get_results = lambda self, *a: None

    g = dict(np.load(os.path.join(GOLDEN, "bprmf_tiny.npz")))
    data = DataSet(_config(10), _frames(g))
    with pytest.raises(Exception):
        M(data, SimpleNamespace(), SimpleNamespace(meta=SimpleNamespace(validation_metric="MAP@10")))
    with pytest.raises(Exception):
        M(data, SimpleNamespace(), SimpleNamespace(meta=SimpleNamespace(validation_metric="nDCG@5")))
    with pytest.raises(Exception):
        M(data, SimpleNamespace(), SimpleNamespace(meta=SimpleNamespace(validation_rate=5), epochs=2))
    m = M(data, SimpleNamespace(), SimpleNamespace(meta=SimpleNamespace(), epochs=3, seed=7))
    assert m.get_base_params_shortcut() == "seed=7_e=3_bs=-1"


def test_neumf_reference_sampler_replays_reference_epochs():
    """custom_sampler.py:14-48 (NeuMF): same (user, item, label) sequence for two epochs, m = 0 and m = 3."""
    from elliot_b200.recommender.neumf import ReferenceSampler
    g = np.load(os.path.join(GOLDEN, "bprmf_tiny.npz")); s = np.load(os.path.join(GOLDEN, "samplers_tiny.npz"))
    nu = len(g["users"])
    i_train = {u: {int(i): 1.0 for i in g["ui_indices"][g["ui_indptr"][u]:g["ui_indptr"][u + 1]]} for u in range(nu)}
    for m in (0, 3):
        smp = ReferenceSampler(i_train, m)
        for ep in range(2):
            u, i, y = smp.epoch()
            want = s[f"neumf_m{m}_ep{ep}"]
            assert np.array_equal(u, want[0]) and np.array_equal(i, want[1]) and np.array_equal(y, want[2]), (m, ep)


def test_multivae_epoch_order_replays_reference():
    """sparse_sampler.py:9-25: per-epoch user permutation from random.seed(42)."""
    import random
    from elliot_b200.recommender.multi_vae import epoch_user_order
    s = np.load(os.path.join(GOLDEN, "samplers_tiny.npz"))
    random.seed(42)
    for ep in range(2):
        assert epoch_user_order(s["vae_perm"].shape[1]) == s["vae_perm"][ep].tolist()


@pytest.mark.parametrize("case", ["tiny", "small"])
def test_mf2020_sampler_replays_reference_epochs(case):
    """RendleSampler (host side of MF2020 exact mode) against the reference's custom_sampler_rendle run: positives in
    sp_i_train.nonzero() order, then the two epoch sample lists, item for item."""
    from elliot_b200.recommender.mf2020 import RendleSampler
    g = dict(np.load(os.path.join(GOLDEN, f"mf2020_{case}.npz")))
    gb = dict(np.load(os.path.join(GOLDEN, f"bprmf_{case}.npz")))
    data = DataSet(_config(int(g["k"])), _frames(gb))
    rs = np.random.RandomState(int(g["seed"]))
    U0 = rs.normal(0, 0.1, (data.num_users, int(g["d"]))); V0 = rs.normal(0, 0.1, (data.num_items, int(g["d"])))
    assert np.array_equal(U0, g["U0"]) and np.array_equal(V0, g["V0"])
    smp = RendleSampler(data.sp_i_train, int(g["m"]), int(g["seed"]), rs)
    assert np.array_equal(np.stack([smp.pos_u, smp.pos_i], 1), g["positives"])
    for ep in range(int(g["epochs"])):
        assert np.array_equal(smp.epoch(), g[f"samples_ep{ep}"])


def test_device_evaluator_tables_match_per_user_loops(golden):
    """Evaluator._device_set (inputs of eb_eval_topk_f64): item-sorted relevant rows, gains aligned with them, IDCG@k
    per user — checked on the CPU against the straightforward per-user computation for every cutoff."""
    import math
    import torch
    g = golden
    data = DataSet(_config(int(g["k"])), _frames(g))
    ev = Evaluator(data, SimpleNamespace(meta=SimpleNamespace()))
    indptr, idx, gain = ev._sets["test"]
    for k in (1, 3, int(g["k"])):
        ip, items, gains, idcg, disc = (t.numpy() for t in ev._device_set("test", k, torch.device("cpu")))
        assert np.array_equal(ip, indptr) and len(disc) == k
        assert np.allclose(disc, [math.log(2) / math.log(r + 2) for r in range(k)], rtol=0, atol=0)
        for u in range(data.num_users):
            lo, hi = indptr[u], indptr[u + 1]
            order = np.argsort(idx[lo:hi], kind="stable")
            assert np.array_equal(items[lo:hi], idx[lo:hi][order]) and np.array_equal(gains[lo:hi], gain[lo:hi][order])
            ideal = np.sort(gain[lo:hi])[::-1][:k]
            assert abs(idcg[u] - float((ideal * disc[:len(ideal)]).sum())) < 1e-12


def test_shard_range_and_owner_of_agree():
    from elliot_b200.parallel import owner_of, shard_range
    for n in (0, 1, 7, 64, 1000, 1001):
        for world in (1, 2, 3, 8):
            edges = [shard_range(n, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == n and all(a[1] == b[0] for a, b in zip(edges, edges[1:]))
            assert max(hi - lo for lo, hi in edges) - min(hi - lo for lo, hi in edges) <= 1
            for row in range(n):
                r = owner_of(row, n, world)
                assert edges[r][0] <= row < edges[r][1]
