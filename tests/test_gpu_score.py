"""GPU parity tests for full-catalogue scoring + mask + top-k through the C ABI."""
import numpy as np
import pytest
import torch

import oracle
from elliot_b200 import ops

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _mask_dev(g):
    return (torch.from_numpy(g["ui_indptr"].astype(np.int64)).to(DEV),
            torch.from_numpy(g["ui_indices"].astype(np.int32)).to(DEV))


def test_topk_f64_equals_reference_lists(golden):
    """fp64 tables: index lists identical to MFModel.get_user_predictions, scores within 1e-12."""
    g = golden
    d, k = int(g["d"]), int(g["k"])
    U = torch.from_numpy(g["U"]).to(DEV); V = torch.from_numpy(g["V"]).to(DEV); b = torch.from_numpy(g["b"]).to(DEV)
    mp, mi = _mask_dev(g)
    idx, val = ops.score_topk(U, V, b, d, k, mp, mi)
    torch.cuda.synchronize()
    assert np.array_equal(idx.cpu().numpy(), g["rec_idx"])
    fin = np.isfinite(g["rec_val"])
    assert np.abs(val.cpu().numpy() - g["rec_val"])[fin].max() < 1e-12


def test_topk_f32_padded_tables(golden):
    """fp32 padded tables vs the oracle run on the SAME fp32-rounded values: identical lists
    wherever the oracle's gap between consecutive scores exceeds 1e-5 (fp32 dot noise)."""
    g = golden
    d, k = int(g["d"]), int(g["k"]); ld = ops.padded_dim(d)
    nu, ni = len(g["users"]), len(g["items"])
    U32 = np.zeros((nu, ld), np.float32); U32[:, :d] = g["U"]
    V32 = np.zeros((ni, ld), np.float32); V32[:, :d] = g["V"]
    b32 = g["b"].astype(np.float32)
    mp, mi = _mask_dev(g)
    idx, val = ops.score_topk(torch.from_numpy(U32).to(DEV), torch.from_numpy(V32).to(DEV),
                              torch.from_numpy(b32).to(DEV), d, k, mp, mi)
    torch.cuda.synchronize()
    oi, ov = oracle.user_topk(U32[:, :d].astype(np.float64), V32[:, :d].astype(np.float64), b32.astype(np.float64),
                              g["ui_indptr"], g["ui_indices"], np.arange(nu), k + 1)
    idx, val = idx.cpu().numpy(), val.cpu().numpy()
    gaps = np.abs(np.diff(ov, axis=1))
    safe = np.all(~np.isfinite(gaps) | (gaps > 1e-5), axis=1)
    assert safe.mean() > 0.9
    assert np.array_equal(idx[safe], oi[safe, :k])
    fin = np.isfinite(ov[:, :k])
    assert np.abs(val - ov[:, :k])[fin].max() < 1e-5


def test_topk_user_subset_and_no_mask(golden_small):
    g = golden_small
    d = int(g["d"])
    U = torch.from_numpy(g["U"]).to(DEV); V = torch.from_numpy(g["V"]).to(DEV)
    users = np.array([5, 0, 399, 17, 17], np.int32)
    idx, val = ops.score_topk(U, V, None, d, 7, users=torch.from_numpy(users).to(DEV))
    oi, ov = oracle.user_topk(g["U"], g["V"], None, None, None, users, 7)
    assert np.array_equal(idx.cpu().numpy(), oi) and np.abs(val.cpu().numpy() - ov).max() < 1e-12


def test_topk_fewer_candidates_than_k():
    """A user who rated all but 3 items: 3 finite entries, then idx -1 / -inf padding; exact ties
    resolve to the lower item index."""
    ni, d, k = 12, 8, 5
    rs = np.random.RandomState(3)
    U = rs.normal(size=(2, d)); V = rs.normal(size=(ni, d))
    V[7] = V[4]                                   # exact tie between items 4 and 7
    indptr = np.array([0, ni - 3, ni - 3], np.int64)
    indices = np.array([x for x in range(ni) if x not in (2, 4, 7)], np.int32)
    idx, val = ops.score_topk(torch.from_numpy(U).to(DEV), torch.from_numpy(V).to(DEV), None, d, k,
                              torch.from_numpy(indptr).to(DEV), torch.from_numpy(indices).to(DEV))
    idx, val = idx.cpu().numpy(), val.cpu().numpy()
    oi, ov = oracle.user_topk(U, V, None, indptr, indices, np.arange(2), k)
    assert np.array_equal(idx, oi)
    assert set(idx[0, :3]) == {2, 4, 7} and list(idx[0, 3:]) == [-1, -1] and np.isinf(val[0, 3:]).all()
    p4, p7 = list(idx[0]).index(4), list(idx[0]).index(7)
    assert p4 + 1 == p7
