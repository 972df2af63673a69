"""Sibling models on the hot path's kernels: MultiDAE and GMF against the fp64 restatements of the
reference's model code (oracle/tf_models.py, TensorFlow parity UNPINNED; their backward passes are cross-checked against
autodiff in tests/test_oracle_autograd.py), plus YAML-driven runs through run_experiment."""
import os

import numpy as np
import pytest
import torch
import yaml

from elliot_b200 import ops
from oracle import tf_models as tfm

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
rel = lambda a, b: np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-12)


def _csr(nu, ni, per, seed):
    rs = np.random.RandomState(seed)
    rows = [np.sort(rs.choice(ni, size=rs.randint(1, 2 * per), replace=False)).astype(np.int32) for _ in range(nu)]
    indptr = np.zeros(nu + 1, np.int64); indptr[1:] = np.cumsum([len(r) for r in rows])
    return rows, torch.from_numpy(indptr).to(DEV), torch.from_numpy(np.concatenate(rows)).to(DEV)


def test_multidae_step_matches_restatement():
    from elliot_b200.recommender.multi_dae import DenoisingAutoEncoder
    nu, ni, H, L, B = 300, 1000, 64, 24, 160
    rows_np, indptr, indices = _csr(nu, ni, 30, 0)
    m = DenoisingAutoEncoder(ni, H, L, 1e-3, 0.0, 0.01, 42, indptr, indices, DEV)
    assert m.P["W2"].shape == (L, H)
    for k in ("b1", "b2", "b3", "b4"):
        m.P[k].normal_(0, 0.05)
    P = {"W1": m.P["W1"].double().cpu().numpy(), "b1": m.P["b1"][:H].double().cpu().numpy(),
         "W2": m.P["W2"].double().cpu().numpy().T, "b2": m.P["b2"][:L].double().cpu().numpy(),
         "W3": m.P["W3"].double().cpu().numpy().T, "b3": m.P["b3"][:H].double().cpu().numpy(),
         "W4": m.P["W4"].double().cpu().numpy().T, "b4": m.P["b4"][:ni].double().cpu().numpy()}
    batch = np.random.RandomState(1).choice(nu, B, replace=False).astype(np.int32)
    X = np.zeros((B, ni))
    for r, u in enumerate(batch):
        X[r, rows_np[u]] = 1.0
    loss_ref, G, logits_ref = tfm.multidae_forward_backward(P, X)
    rows = torch.from_numpy(batch).to(DEV)
    _, _, _, _, logits = m._forward(rows, 1)
    assert rel(logits.double().cpu().numpy(), logits_ref) < 2e-2
    m._refresh()
    loss = m.train_step(rows, 0.0)
    assert abs(loss - loss_ref) < 2e-2 * abs(loss_ref)
    for k, tr in (("W4", True), ("W3", True), ("W2", True), ("W1", False)):       # first Adam step: m = 0.1 g
        g_ref = G[k].T if tr else G[k]
        assert rel(m.M[k].double().cpu().numpy() / 0.1, g_ref) < 5e-2, k
    for k, n in (("b4", ni), ("b3", H), ("b2", L), ("b1", H)):
        assert rel(m.M[k][:n].double().cpu().numpy() / 0.1, G[k]) < 5e-2, k
    idx, val = m.predict_topk(torch.arange(0, 64, dtype=torch.int32, device=DEV), 10, indptr, indices)
    idx = idx.cpu().numpy()
    for r in range(64):
        assert not set(idx[r]) & set(rows_np[r].tolist())


def test_multidae_wiring_is_exact_with_the_fp32_checking_gemm():
    from elliot_b200.recommender.multi_dae import DenoisingAutoEncoder
    nu, ni, H, L, B = 300, 1000, 64, 24, 160
    rows_np, indptr, indices = _csr(nu, ni, 30, 0)
    with ops.exact_gemm():
        m = DenoisingAutoEncoder(ni, H, L, 1e-3, 0.0, 0.01, 42, indptr, indices, DEV)
        for k in ("b1", "b2", "b3", "b4"):
            m.P[k].normal_(0, 0.05)
        m._refresh()
        P = {"W1": m.P["W1"].double().cpu().numpy(), "b1": m.P["b1"][:H].double().cpu().numpy(),
             "W2": m.P["W2"].double().cpu().numpy().T, "b2": m.P["b2"][:L].double().cpu().numpy(),
             "W3": m.P["W3"].double().cpu().numpy().T, "b3": m.P["b3"][:H].double().cpu().numpy(),
             "W4": m.P["W4"].double().cpu().numpy().T, "b4": m.P["b4"][:ni].double().cpu().numpy()}
        batch = np.random.RandomState(1).choice(nu, B, replace=False).astype(np.int32)
        X = np.zeros((B, ni))
        for r, u in enumerate(batch):
            X[r, rows_np[u]] = 1.0
        loss_ref, G, logits_ref = tfm.multidae_forward_backward(P, X)
        loss = m.train_step(torch.from_numpy(batch).to(DEV), 0.0)
    assert abs(loss - loss_ref) < 1e-5 * abs(loss_ref)
    for k, tr in (("W4", True), ("W3", True), ("W2", True), ("W1", False)):
        assert rel(m.M[k].double().cpu().numpy() / 0.1, G[k].T if tr else G[k]) < 1e-4, k
    for k, n in (("b4", ni), ("b3", H), ("b2", L), ("b1", H)):
        assert rel(m.M[k][:n].double().cpu().numpy() / 0.1, G[k]) < 1e-4, k


def test_gmf_fused_step_matches_restatement():
    """fp32 CUDA-core kernel vs the fp64 restatement: loss 1e-6, every gradient 1e-5 relative (no bf16 anywhere)."""
    from elliot_b200.recommender.gmf import GeneralizedMatrixFactorizationModel
    nu, ni, f, B = 400, 300, 10, 5000                       # f = 10: the reference default, padded to 12 columns / stride 16
    m = GeneralizedMatrixFactorizationModel(nu, ni, f, True, 1e-3, 42, DEV)
    assert not m.P["U"][:, f:].any() and not m.P["h"][f:].any()
    gen = torch.Generator(device=DEV); gen.manual_seed(1)
    u = torch.randint(0, nu, (B,), device=DEV, generator=gen, dtype=torch.int32)
    it = torch.randint(0, ni, (B,), device=DEV, generator=gen, dtype=torch.int32)
    y = (torch.rand(B, device=DEV, generator=gen) < 0.5).float()
    P = {"U": m.P["U"][:, :f].double().cpu().numpy(), "I": m.P["I"][:, :f].double().cpu().numpy(), "h": m.P["h"][:f].double().cpu().numpy()}
    loss_ref, G, _ = tfm.gmf_forward_backward(P, u.cpu().numpy(), it.cpu().numpy(), y.double().cpu().numpy())
    loss = torch.zeros(1, dtype=torch.float64, device=DEV)
    ops.gmf_step_grads(m.P["U"], m.P["I"], m.P["h"], m.fp, u, it, y, m.G["U"], m.G["I"], m.G["h"], loss=loss)
    assert abs(loss.item() - loss_ref) < 1e-6 * abs(loss_ref)
    assert rel(m.G["U"][:, :f].double().cpu().numpy(), G["U"]) < 1e-5 and rel(m.G["I"][:, :f].double().cpu().numpy(), G["I"]) < 1e-5
    assert rel(m.G["h"][:f].double().cpu().numpy(), G["h"]) < 1e-5
    assert not m.G["U"][:, f:].any() and not m.G["h"][f:].any()          # padding columns never receive gradient
    for k in m.G:
        m.G[k].zero_()
    # one full train step = the restatement's Keras Adam step
    opt = tfm.KerasAdam(1e-3); opt.begin_step()
    want = {k: v.copy() for k, v in P.items()}
    for k in want:
        opt.apply(k, want[k], G[k])
    m.train_step((u, it, y))
    assert np.abs(m.P["U"][:, :f].double().cpu().numpy() - want["U"]).max() < 2e-6
    assert np.abs(m.P["h"][:f].double().cpu().numpy() - want["h"]).max() < 2e-6
    # scoring: probabilities of the masked top-k equal sigmoid of the restated logits
    indptr = torch.zeros(nu + 1, dtype=torch.int64, device=DEV); indices = torch.zeros(1, dtype=torch.int32, device=DEV)
    idx, val = m.get_recs_topk(5, indptr, indices)
    Pn = {"U": m.P["U"][:, :f].double().cpu().numpy(), "I": m.P["I"][:, :f].double().cpu().numpy(), "h": m.P["h"][:f].double().cpu().numpy()}
    full = 1 / (1 + np.exp(-((Pn["U"] * Pn["h"]) @ Pn["I"].T)))
    top = np.sort(full, 1)[:, ::-1][:, :5]
    assert np.abs(val.double().cpu().numpy() - top).max() < 1e-5


def test_pointwise_sampler_distribution(golden_small):
    g = golden_small
    nu, ni = len(g["users"]), len(g["items"])
    indptr = torch.from_numpy(g["ui_indptr"].astype(np.int64)).to(DEV)
    srt = g["ui_indices"].astype(np.int32).copy()
    for u in range(nu):
        srt[g["ui_indptr"][u]:g["ui_indptr"][u + 1]].sort()
    srt = torch.from_numpy(srt).to(DEV)
    n = 200000
    u, i, y = (t.cpu().numpy() for t in ops.pointwise_sample_philox(nu, ni, indptr, srt, n, 42))
    rows = [set(g["ui_indices"][g["ui_indptr"][x]:g["ui_indptr"][x + 1]].tolist()) for x in range(nu)]
    assert abs(y.mean() - 0.5) < 0.01                                    # a fair bit per sample (pointwise_pos_neg_sampler.py:39)
    for t in range(0, n, 41):
        assert (i[t] in rows[u[t]]) == (y[t] == 1.0)
    cnt = np.bincount(u, minlength=nu); chi = ((cnt - n / nu) ** 2 / (n / nu)).sum()
    assert abs(chi - (nu - 1)) < 6 * np.sqrt(2 * (nu - 1))


@pytest.mark.parametrize("key,block", [
    ("MultiDAE", {"epochs": 3, "batch_size": 64, "intermediate_dim": 64, "latent_dim": 32, "lr": 0.003, "dropout_pkeep": 0.8}),
    ("GMF", {"epochs": 4, "batch_size": 512, "mf_factors": 16, "lr": 0.01}),
])
def test_sibling_models_train_from_yaml(tmp_path, golden_small, key, block):
    from elliot_b200 import run_experiment
    g = golden_small
    for name in ("train", "test"):
        with open(tmp_path / f"{name}.tsv", "w") as fh:
            for u, i, r in g[name]:
                fh.write(f"{int(u)}\t{int(i)}\t{r}\n")
    cfg = {"experiment": {"dataset": "golden", "data_config": {"strategy": "fixed", "train_path": "train.tsv", "test_path": "test.tsv"},
                          "top_k": 10, "evaluation": {"simple_metrics": ["nDCG", "HR"]},
                          "path_output_rec_result": "out/recs", "path_output_rec_weight": "out/weights",
                          "path_output_rec_performance": "out/perf",
                          "models": {key: {"meta": {"save_recs": False}, "seed": 7, **block}}}}
    p = tmp_path / "cfg.yml"; p.write_text(yaml.safe_dump(cfg))
    res = run_experiment(str(p))[0]
    hist = [h[10]["nDCG"] for h in res["history"]]
    assert len(hist) == block["epochs"] and all(np.isfinite(hist)) and hist[-1] > 0.0
    assert res["name"].startswith("MultiDAE_" if key == "MultiDAE" else "GeneralizedMF_")
    # training moves the ranking away from the untrained model's
    assert hist[-1] != hist[0]
