"""MF2020 (pointwise logistic MF) through the C ABI: the exact fp64 kernel against the REFERENCE's
own MF2020 run (tests/golden/mf2020_*.npz, minted by oracle/gen_golden.py) and against the oracle on adversarial
sample lists; the YAML-driven model; the fp32 throughput kernel (distribution + learning)."""
import os

import numpy as np
import pytest
import torch
import yaml

import oracle
from elliot_b200 import ops

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _dev(a, dt):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV, dt)


@pytest.mark.parametrize("case", ["tiny", "small"])
def test_exact_kernel_reproduces_reference_run(case):
    g = dict(np.load(os.path.join(GOLDEN, f"mf2020_{case}.npz")))
    d = int(g["d"])
    U, V = _dev(g["U0"], torch.float64), _dev(g["V0"], torch.float64)
    ub = torch.zeros(U.shape[0], dtype=torch.float64, device=DEV); ib = torch.zeros(V.shape[0], dtype=torch.float64, device=DEV)
    gb = torch.zeros(1, dtype=torch.float64, device=DEV)
    for ep in range(int(g["epochs"])):
        s = g[f"samples_ep{ep}"]
        bl = torch.zeros((len(s) + 99999) // 100000, dtype=torch.float64, device=DEV)
        ops.mf_pointwise_exact_f64(U, V, ub, ib, gb, d, _dev(s[:, 0], torch.int32), _dev(s[:, 1], torch.int32),
                                   _dev(s[:, 2], torch.int32), float(g["lr"]), float(g["reg"]), batch_loss=bl)
        assert np.allclose(bl.cpu().numpy(), g[f"batch_loss_ep{ep}"], rtol=1e-11, atol=0)
    assert np.abs(U.cpu().numpy() - g["U"]).max() < 1e-12 and np.abs(V.cpu().numpy() - g["V"]).max() < 1e-12
    assert np.abs(ub.cpu().numpy() - g["ub"]).max() < 1e-12 and np.abs(ib.cpu().numpy() - g["ib"]).max() < 1e-12
    assert abs(gb.item() - float(g["gb"])) < 1e-12


@pytest.mark.parametrize("d,n_u,n_i,n", [(1, 3, 2, 500), (7, 5, 4, 2000), (33, 40, 30, 3000), (64, 2, 2, 1000),
                                         (100, 50, 60, 2500), (200, 9, 9, 700)])
def test_exact_kernel_equals_oracle_on_conflict_heavy_lists(d, n_u, n_i, n):
    """Tiny id ranges: consecutive samples keep hitting the row the previous one rewrote (register hand-over path),
    every row width class (d <= 32 / 64 / 128 / 256), ragged tail batches."""
    rs = np.random.RandomState(d)
    U0 = rs.normal(0, 0.1, (n_u, d)); V0 = rs.normal(0, 0.1, (n_i, d))
    su = rs.randint(n_u, size=n).astype(np.int32); si = rs.randint(n_i, size=n).astype(np.int32)
    sr = rs.randint(2, size=n).astype(np.int32)
    su[10:20] = su[10]; si[30:45] = si[30]; su[50:60] = su[50]; si[50:60] = si[50]          # runs of identical rows
    U, V, ub, ib = U0.copy(), V0.copy(), rs.normal(0, 0.01, n_u), rs.normal(0, 0.01, n_i)
    dU, dV, dub, dib = (_dev(x, torch.float64) for x in (U, V, ub, ib))
    dgb = torch.tensor([0.05], dtype=torch.float64, device=DEV)
    batch = 300
    want_gb, want_bl = oracle.mf2020_update_seq(U, V, ub, ib, 0.05, su, si, sr, 0.05, 0.01, batch=batch)
    bl = torch.zeros((n + batch - 1) // batch, dtype=torch.float64, device=DEV)
    ops.mf_pointwise_exact_f64(dU, dV, dub, dib, dgb, d, _dev(su, torch.int32), _dev(si, torch.int32), _dev(sr, torch.int32),
                               0.05, 0.01, batch=batch, batch_loss=bl)
    assert np.abs(dU.cpu().numpy() - U).max() < 1e-12 and np.abs(dV.cpu().numpy() - V).max() < 1e-12
    assert np.abs(dub.cpu().numpy() - ub).max() < 1e-12 and np.abs(dib.cpu().numpy() - ib).max() < 1e-12
    assert abs(dgb.item() - want_gb) < 1e-12 and np.allclose(bl.cpu().numpy(), want_bl, rtol=1e-11)


def test_exact_kernel_empty_list_is_a_no_op():
    U = torch.ones(2, 4, dtype=torch.float64, device=DEV); V = U.clone()
    z = torch.zeros(2, dtype=torch.float64, device=DEV); gb = torch.zeros(1, dtype=torch.float64, device=DEV)
    e = torch.empty(0, dtype=torch.int32, device=DEV)
    ops.mf_pointwise_exact_f64(U, V, z, z.clone(), gb, 4, e, e, e, 0.1, 0.0)
    assert bool((U == 1).all()) and gb.item() == 0.0


def _write_case(tmp_path, g, gm, extra=None):
    for name, a in (("train", g["train"]), ("test", g["test"])):
        with open(tmp_path / f"{name}.tsv", "w") as f:
            for u, i, r in a:
                f.write(f"{int(u)}\t{int(i)}\t{r}\n")
    # validation_rate = epochs: only the final epoch is evaluated, which is what the golden metrics describe
    block = {"meta": {"save_recs": False, "save_weights": False, "validation_rate": int(gm["epochs"])}, "epochs": int(gm["epochs"]), "factors": int(gm["d"]),
             "seed": int(gm["seed"]), "lr": float(gm["lr"]), "reg": float(gm["reg"]), "m": int(gm["m"])}
    block.update(extra or {})
    cfg = {"experiment": {"dataset": "golden", "data_config": {"strategy": "fixed", "train_path": "train.tsv", "test_path": "test.tsv"},
                          "top_k": int(gm["k"]), "evaluation": {"simple_metrics": ["nDCG", "HR", "Precision", "Recall"]},
                          "path_output_rec_result": "out/recs", "path_output_rec_weight": "out/weights",
                          "path_output_rec_performance": "out/perf", "models": {"MF2020": block}}}
    p = tmp_path / "cfg.yml"
    p.write_text(yaml.safe_dump(cfg))
    return str(p)


@pytest.mark.parametrize("case,ev", [("tiny", "host"), ("small", "host"), ("small", "device")])
def test_yaml_driven_mf2020_equals_reference_run(tmp_path, case, ev):
    """run_experiment on the YAML block of MF.py:41-52: the reference run's metrics (reference Evaluator) to 1e-12
    and its top-k lists item for item."""
    from elliot_b200 import run_experiment
    g = dict(np.load(os.path.join(GOLDEN, f"bprmf_{case}.npz"))); gm = dict(np.load(os.path.join(GOLDEN, f"mf2020_{case}.npz")))
    res = run_experiment(_write_case(tmp_path, g, gm, {"b200_eval": ev}))
    k = int(gm["k"])
    want = dict(zip(gm["metric_names"].tolist(), gm["metric_vals"].tolist()))
    for m in want:
        assert abs(res[0]["test_results"][k][m] - want[m]) < 1e-12, (m, res[0]["test_results"][k][m], want[m])


@pytest.mark.parametrize("case", ["tiny", "small"])
def test_model_lists_and_epoch_loss_equal_reference(case):
    from types import SimpleNamespace
    import pandas as pd
    from elliot_b200.dataset import DataSet
    from elliot_b200.recommender import MF2020
    g = dict(np.load(os.path.join(GOLDEN, f"bprmf_{case}.npz"))); gm = dict(np.load(os.path.join(GOLDEN, f"mf2020_{case}.npz")))
    k = int(gm["k"])
    f = lambda a: pd.DataFrame({"userId": a[:, 0].astype(np.int64), "itemId": a[:, 1].astype(np.int64), "rating": a[:, 2]})
    cfg = SimpleNamespace(config_test=False, top_k=k, path_output_rec_result="/tmp/eb_mf/recs", path_output_rec_weight="/tmp/eb_mf/w",
                          path_output_rec_performance="/tmp/eb_mf/p", dataset="golden",
                          evaluation=SimpleNamespace(simple_metrics=["nDCG"], relevance_threshold=0, paired_ttest=False, cutoffs=[k]))
    data = DataSet(cfg, (f(g["train"]), f(g["test"])))
    params = SimpleNamespace(meta=SimpleNamespace(save_recs=False, save_weights=False), epochs=int(gm["epochs"]),
                             factors=int(gm["d"]), seed=int(gm["seed"]), lr=float(gm["lr"]), reg=float(gm["reg"]), m=int(gm["m"]))
    model = MF2020(data=data, config=cfg, params=params)
    model.train()
    assert np.allclose([l * (e + 1) for e, l in enumerate(model._losses)], gm["epoch_loss"], rtol=1e-11)
    idx, val = model.get_recommendations_tensors(k)
    assert np.array_equal(idx.cpu().numpy(), gm["rec_idx"])
    ok = gm["rec_idx"] >= 0
    assert np.abs(val.cpu().numpy()[ok] - gm["rec_val"][ok]).max() < 1e-12
    st = model._model.get_model_state()
    assert set(st) == {"_global_bias", "_user_bias", "_item_bias", "_user_factors", "_item_factors"}     # MF_model.py:150-158
    assert abs(st["_global_bias"] - float(gm["gb"])) < 1e-12 and np.abs(st["_user_factors"] - gm["U"]).max() < 1e-12


def test_throughput_kernel_visits_every_sample_once_and_learns():
    g = dict(np.load(os.path.join(GOLDEN, "mf2020_small.npz")))
    pos = g["positives"]; m, d = 3, 64
    n_u, n_i = g["U0"].shape[0], g["V0"].shape[0]
    ld = ops.padded_dim(d)
    U = torch.zeros(n_u, ld, device=DEV); U[:, :d] = _dev(g["U0"], torch.float32)
    V = torch.zeros(n_i, ld, device=DEV); V[:, :d] = _dev(g["V0"], torch.float32)
    ub = torch.zeros(n_u, device=DEV); ib = torch.zeros(n_i, device=DEV); gb = torch.zeros(1, device=DEV)
    pu, pi = _dev(pos[:, 0], torch.int32), _dev(pos[:, 1], torch.int32)
    n = len(pos) * (1 + m)
    out = tuple(torch.empty(n, dtype=torch.int32, device=DEV) for _ in range(3))
    losses = []
    for ep in range(6):
        loss = torch.zeros(1, dtype=torch.float64, device=DEV)
        for first in range(0, n, 1024):                               # small launches: few stale hits per row
            ops.mf_pointwise_step_f32(U, V, ub, ib, gb, d, pu, pi, m, n_i, 11, ep, 0.05, 0.005, loss=loss,
                                      out=out if ep < 2 else None, first=first, count=min(1024, n - first))
        losses.append(loss.item() / n)
        if ep < 2:
            ou, oi, orr = (x.cpu().numpy() for x in out)
            # every positive exactly once, m label-0 samples per positive with the positive's user, items in range
            p_rows = np.stack([ou[orr == 1], oi[orr == 1]], 1)
            assert len(p_rows) == len(pos)
            assert np.array_equal(p_rows[np.lexsort((p_rows[:, 1], p_rows[:, 0]))], pos[np.lexsort((pos[:, 1], pos[:, 0]))])
            assert (orr == 0).sum() == m * len(pos) and oi.min() >= 0 and oi.max() < n_i
            assert np.array_equal(np.bincount(ou[orr == 0], minlength=n_u), m * np.bincount(pos[:, 0], minlength=n_u))
            # negatives are uniform over items: chi-square-ish bound on the histogram
            h = np.bincount(oi[orr == 0], minlength=n_i); e = h.sum() / n_i
            assert abs(h - e).max() < 6 * np.sqrt(e) + 6
            if ep == 1:
                assert not np.array_equal(first_order, ou)                       # a new visiting order every epoch
            first_order = ou.copy()
    assert bool(torch.isfinite(U).all()) and losses[-1] < 0.95 * losses[0], losses
    # same trend as the oracle run from the same start (sequential fp64, reference order): final mean loss within 10 %
    import random
    rs = np.random.RandomState(3); pr = random.Random(3)
    Uo, Vo = g["U0"].copy(), g["V0"].copy(); ubo, ibo, gbo = np.zeros(n_u), np.zeros(n_i), 0.0
    for ep in range(6):
        s = oracle.mf2020_epoch_samples(rs, pr, pos[:, 0], pos[:, 1], n_i, m)
        gbo, bl = oracle.mf2020_update_seq(Uo, Vo, ubo, ibo, gbo, s[:, 0], s[:, 1], s[:, 2], 0.05, 0.005)
    assert abs(losses[-1] - bl.sum() / n) < 0.1 * bl.sum() / n, (losses, bl.sum() / n)
    assert abs(gb.item() - gbo) < 0.15 * abs(gbo) + 0.02, (gb.item(), gbo)


def test_global_bias_closed_form_is_stable_for_one_huge_launch():
    """A whole epoch in ONE launch: rows of this small case take many stale hits (Hogwild noise), but the global
    bias — hit by every sample — must land near the launch's fixed point instead of overshooting by lr * n."""
    g = dict(np.load(os.path.join(GOLDEN, "mf2020_small.npz")))
    pos = g["positives"]; m, d = 3, 64
    n_u, n_i = g["U0"].shape[0], g["V0"].shape[0]
    ld = ops.padded_dim(d)
    U = torch.zeros(n_u, ld, device=DEV); V = torch.zeros(n_i, ld, device=DEV)          # zero factors: pred = gb only
    ub = torch.zeros(n_u, device=DEV); ib = torch.zeros(n_i, device=DEV); gb = torch.zeros(1, device=DEV)
    ops.mf_pointwise_step_f32(U, V, ub, ib, gb, d, _dev(pos[:, 0], torch.int32), _dev(pos[:, 1], torch.int32), m, n_i, 5, 0,
                              0.0, 0.0)                                              # lr = 0: nothing moves
    assert gb.item() == 0.0
    # with only the global bias learning (rows frozen by lr=0 is not expressible, so compare against the fixed point):
    # labels are 1 : m, so sigmoid(gb*) = 1/(1+m)  ->  gb* = -log(m); one launch from 0 must move towards it, not past it
    ops.mf_pointwise_step_f32(U, V, ub, ib, gb, d, _dev(pos[:, 0], torch.int32), _dev(pos[:, 1], torch.int32), m, n_i, 5, 1,
                              0.05, 0.0)
    assert -np.log(m) - 0.35 < gb.item() < 0.0
