"""Device evaluation: eb_eval_topk_f64 against (1) the numbers the REFERENCE Evaluator produced on
the reference's own top-k lists (tests/golden, minted by oracle/gen_golden.py) and (2) the host evaluator mirror on
random lists with graded gains, ragged / empty test rows, empty slots and test-only items."""
import os
from types import SimpleNamespace

import numpy as np
import pandas as pd
import pytest
import torch

from elliot_b200 import ops
from elliot_b200.dataset import DataSet
from elliot_b200.evaluation import Evaluator

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _frames(g):
    f = lambda a: pd.DataFrame({"userId": a[:, 0].astype(np.int64), "itemId": a[:, 1].astype(np.int64), "rating": a[:, 2]})
    return f(g["train"]), f(g["test"])


def _config(k, cutoffs=None, thr=0):
    return SimpleNamespace(config_test=False, top_k=k,
                           evaluation=SimpleNamespace(simple_metrics=["nDCG", "HR", "Precision", "Recall"],
                                                      relevance_threshold=thr, paired_ttest=False, cutoffs=cutoffs or [k]))


def test_device_metrics_equal_reference_evaluator(golden):
    g = golden
    k = int(g["k"])
    data = DataSet(_config(k), _frames(g))
    ev = Evaluator(data, SimpleNamespace(meta=SimpleNamespace()))
    idx = torch.from_numpy(g["rec_idx"].astype(np.int32)).to(DEV)
    got = ev.eval_tensors(idx)[k]["test_results"]
    want = dict(zip(g["metric_names"].tolist(), g["metric_vals"].tolist()))
    for m in want:
        assert abs(got[m] - want[m]) < 1e-12, (m, got[m], want[m])


@pytest.mark.parametrize("k,top_k", [(1, 1), (5, 10), (10, 10), (16, 20), (50, 50)])
def test_device_metrics_equal_host_mirror_on_random_lists(k, top_k):
    rs = np.random.RandomState(k)
    n_users, n_items = 700, 300
    tr = [(u, i, 1.0) for u in range(n_users) for i in rs.choice(n_items, 3, replace=False)]
    te = []
    for u in range(n_users):
        m = 0 if u % 7 == 0 else rs.randint(1, 40)                  # some users have no test items at all
        for i in rs.choice(n_items + 20, m, replace=False):         # ids >= n_items never occur in training
            te.append((u, i, float(rs.randint(1, 6))))
    f = lambda a: pd.DataFrame({"userId": [x[0] for x in a], "itemId": [x[1] for x in a], "rating": [x[2] for x in a]})
    data = DataSet(_config(top_k, cutoffs=[k], thr=2), (f(tr), f(te)))
    ev = Evaluator(data, SimpleNamespace(meta=SimpleNamespace()))
    idx = np.stack([rs.permutation(data.num_items)[:top_k] for _ in range(data.num_users)]).astype(np.int32)
    idx[rs.rand(*idx.shape) < 0.05] = -1                            # empty slots anywhere in the list
    want = ev.eval_arrays(np.arange(data.num_users), idx.astype(np.int64), ev._sets["test"], k)
    got = ev.eval_tensors(torch.from_numpy(idx).to(DEV))[k]["test_results"]
    for m in want:
        assert abs(got[m] - want[m]) < 1e-12, (m, got[m], want[m])
    # per-user values + explicit user ids (a shuffled subset of rows)
    sel = rs.permutation(data.num_users)[:200].astype(np.int32)
    ds = ev._device_set("test", k, torch.device(DEV))
    sums, per = ops.eval_topk(torch.from_numpy(idx[sel]).to(DEV), k, *ds, users=torch.from_numpy(sel).to(DEV), per_user=True)
    per = per.cpu().numpy(); sums = sums.cpu().numpy()
    ok = ~np.isnan(per[:, 0])
    assert sums[0] == ok.sum()
    assert np.allclose(per[ok].sum(0), sums[1:], rtol=1e-12)
    sub = ev.eval_arrays(sel.astype(np.int64), idx[sel].astype(np.int64), ev._sets["test"], k)
    for j, m in enumerate(("nDCG", "HR", "Precision", "Recall")):
        assert abs(sums[1 + j] / sums[0] - sub[m]) < 1e-12
    # deterministic: same bits on a second launch
    sums2, _ = ops.eval_topk(torch.from_numpy(idx[sel]).to(DEV), k, *ds, users=torch.from_numpy(sel).to(DEV))
    assert np.array_equal(sums2.cpu().numpy(), sums)


def test_empty_and_all_skipped():
    z = lambda n, dt: torch.zeros(n, dtype=dt, device=DEV)
    out, _ = ops.eval_topk(torch.empty(0, 10, dtype=torch.int32, device=DEV), 10, z(1, torch.int64), z(1, torch.int32),
                           z(1, torch.float64), z(1, torch.float64), torch.ones(10, dtype=torch.float64, device=DEV))
    assert out.cpu().tolist() == [0.0] * 5
    idx = torch.zeros(5, 10, dtype=torch.int32, device=DEV)
    out, per = ops.eval_topk(idx, 10, z(6, torch.int64), z(1, torch.int32), z(1, torch.float64), z(5, torch.float64),
                             torch.ones(10, dtype=torch.float64, device=DEV), per_user=True)
    assert out.cpu().tolist() == [0.0] * 5 and bool(torch.isnan(per).all())
