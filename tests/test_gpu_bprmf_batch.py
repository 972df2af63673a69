"""BPRMF_batch kernels (gradient + Keras-Adam) against the fp64 restatement in oracle/tf_models.py
(parity with TensorFlow itself is UNPINNED: TF 2.3.2 cannot be run here)."""
import numpy as np
import pytest
import torch

from elliot_b200 import ops
from oracle import tf_models as tfm

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("d", [10, 64])
def test_batch_grad_and_adam_match_restatement(d):
    rs = np.random.RandomState(d)
    nu, ni, B = 50, 40, 600                                   # many duplicate rows per batch
    ld = ops.padded_dim(d)
    Gu, Gi = tfm.glorot_uniform(rs, (nu, d)), tfm.glorot_uniform(rs, (ni, d))
    Bi = rs.normal(0, 0.01, ni)
    l_w, l_b, lr = 0.1, 0.001, 0.001
    nb = (ni + 3) // 4 * 4

    def dev(a, cols=None):
        t = torch.zeros((a.shape[0], ld), device=DEV); t[:, :a.shape[1]] = torch.from_numpy(a).float().to(DEV); return t
    Gud, Gid = dev(Gu), dev(Gi)
    Bid = torch.zeros(nb, device=DEV); Bid[:ni] = torch.from_numpy(Bi).float().to(DEV)
    st = {k: (torch.zeros_like(v), torch.zeros_like(v), torch.zeros_like(v)) for k, v in (("Gu", Gud), ("Gi", Gid), ("Bi", Bid))}
    opt = tfm.KerasAdam(lr)
    loss_d = torch.zeros(1, dtype=torch.float64, device=DEV)
    for step in range(1, 4):
        u = rs.randint(0, nu, B).astype(np.int32); i = rs.randint(0, ni, B).astype(np.int32)
        j = ((i + 1 + rs.randint(0, ni - 1, B)) % ni).astype(np.int32)
        if step == 2:
            Gu[u[0]] *= 300; Gud[u[0]] *= 300                 # drive one diff far below -80: clipped, zero gradient
        loss_ref = tfm.bprmf_batch_step(Gu, Gi, Bi, opt, u, i, j, l_w, l_b)
        loss_d.zero_()
        ops.bpr_batch_grad_f32(Gud, Gid, Bid, st["Gu"][0], st["Gi"][0], st["Bi"][0], d, torch.from_numpy(u).to(DEV),
                               torch.from_numpy(i).to(DEV), torch.from_numpy(j).to(DEV), l_w, l_b, loss=loss_d)
        for name, var in (("Bi", Bid), ("Gu", Gud), ("Gi", Gid)):
            g, m, v = st[name]
            ops.adam_dense_f32(var, m, v, g, lr, step)
        torch.cuda.synchronize()
        assert abs(loss_d.item() - loss_ref) < 2e-4 * abs(loss_ref), (step, loss_d.item(), loss_ref)
        # fp32 vs fp64: Adam's m/(sqrt(v)+eps) normalises the step to ~lr, so 5e-5 abs covers rounding
        assert np.abs(Gud.cpu().numpy()[:, :d] - Gu).max() < 5e-5 * max(1.0, np.abs(Gu).max())
        assert np.abs(Gid.cpu().numpy()[:, :d] - Gi).max() < 5e-5
        assert np.abs(Bid.cpu().numpy()[:ni] - Bi).max() < 5e-5
        assert not st["Gu"][0].any() and not Gud[:, d:].any()        # gradient cleared, padding untouched


def test_bprmf_batch_model_class_trains(tmp_path):
    """YAML-driven BPRMF_batch: loss decreases, nDCG beats the untrained tables."""
    import os, yaml
    from elliot_b200 import run_experiment
    g = dict(np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "bprmf_small.npz")))
    for name in ("train", "test"):
        with open(tmp_path / f"{name}.tsv", "w") as f:
            for u, i, r in g[name]:
                f.write(f"{int(u)}\t{int(i)}\t{r}\n")

    def cfg(epochs, lr):
        c = {"experiment": {"dataset": "golden", "data_config": {"strategy": "fixed", "train_path": "train.tsv", "test_path": "test.tsv"},
                            "top_k": 10, "evaluation": {"simple_metrics": ["nDCG"]},
                            "path_output_rec_result": "o/r", "path_output_rec_weight": "o/w", "path_output_rec_performance": "o/p",
                            "models": {"BPRMF_batch": {"meta": {}, "epochs": epochs, "batch_size": 512, "factors": 32, "lr": lr,
                                                       "l_w": 0.001, "l_b": 0.0}}}}
        p = tmp_path / "c.yml"; p.write_text(yaml.safe_dump(c)); return str(p)
    r0 = run_experiment(cfg(1, 0.0))
    r1 = run_experiment(cfg(25, 0.01))
    assert r1[0]["test_results"][10]["nDCG"] > 1.5 * r0[0]["test_results"][10]["nDCG"] + 0.02
