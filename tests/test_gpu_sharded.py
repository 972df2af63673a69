"""Row-sharded item tables: kernels + the fetch/update/push protocol at world size 1 against the
replicated-table kernel and the oracle; the multi-rank exchange itself is covered by tests/test_parallel_cpu.py
(gloo) and tools/sharded_check.py (torchrun on 2+ GPUs, results in profiles/)."""
import numpy as np
import pytest
import torch

import oracle
from elliot_b200 import ops
from elliot_b200.parallel import ShardedTable, sharded_bpr_step

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def test_gather_scatter_rows():
    g = torch.Generator(device=DEV); g.manual_seed(0)
    T = torch.randn(50, 16, device=DEV, generator=g)
    ids = torch.tensor([3, 3, 49, 0, 17], dtype=torch.int32, device=DEV)
    rows = ops.gather_rows_f32(T, ids)
    assert torch.equal(rows, T[ids.long()])
    T2 = T.clone()
    ops.scatter_add_rows_f32(T2, ids, torch.ones_like(rows))
    want = T.clone(); want.index_add_(0, ids.long(), torch.ones_like(rows))
    assert torch.allclose(T2, want)


@pytest.mark.parametrize("d,with_bias", [(60, True), (64, False), (10, True)])
def test_sharded_step_equals_sequential_on_conflict_free_batch(golden_small, d, with_bias):
    g = golden_small
    rs = np.random.RandomState(d)
    nu, ni = len(g["users"]), len(g["items"])
    ld = ops.padded_dim(d + (1 if with_bias else 0))
    U0 = rs.normal(0, 0.1, (nu, d)); V0 = rs.normal(0, 0.1, (ni, d)); b0 = rs.normal(0, 0.05, ni) if with_bias else np.zeros(ni)
    seen_u, seen_i, keep = set(), set(), []
    for t in range(len(g["tu"])):
        u, i, j = g["tu"][t], g["ti"][t], g["tj"][t]
        if u in seen_u or i in seen_i or j in seen_i: continue
        seen_u.add(u); seen_i.add(i); seen_i.add(j); keep.append(t)
    keep = np.array(keep); tu, ti, tj = g["tu"][keep], g["ti"][keep], g["tj"][keep]
    hp = (0.05, 0.0025, 0.01 if with_bias else 0.0, 0.0025, 0.00025)
    U, V, b = U0.copy(), V0.copy(), b0.copy()
    oracle.bpr_update_seq(U, V, b, tu, ti, tj, *hp)
    Ud = torch.zeros((nu, ld), device=DEV); Ud[:, :d] = torch.from_numpy(U0).float().to(DEV)
    Vd = torch.zeros((ni, ld), device=DEV); Vd[:, :d] = torch.from_numpy(V0).float().to(DEV)
    if with_bias: Vd[:, d] = torch.from_numpy(b0).float().to(DEV)
    items = ShardedTable(ni, Vd)
    loss = torch.zeros(1, dtype=torch.float64, device=DEV)
    sharded_bpr_step(Ud, items, torch.from_numpy(tu).to(DEV), torch.from_numpy(ti).to(DEV), torch.from_numpy(tj).to(DEV), hp,
                     bias_col=d if with_bias else -1, loss=loss)
    torch.cuda.synchronize()
    assert np.abs(Ud.cpu().numpy()[:, :d] - U).max() < 2e-6
    assert np.abs(Vd.cpu().numpy()[:, :d] - V).max() < 2e-6
    if with_bias:
        assert np.abs(Vd.cpu().numpy()[:, d] - b).max() < 2e-6
        assert not Ud[:, d:].any()                                # the bias column never leaks into the user rows
    assert loss.item() > 0
