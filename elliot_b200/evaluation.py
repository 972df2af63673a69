"""Minimal evaluator for the host mirror: nDCG / HR / Precision / Recall with the reference's
definitions (elliot/evaluation/evaluator.py:79-147; ndcg.py:68-125; relevance.py:55,80-82;
hit_rate.py, precision.py, recall.py), vectorised over (users x k) index arrays.

The reference Evaluator is the parity judge and is NOT re-implemented in breadth (SURVEY.md §2
#23 out of scope): tests/test_host_parity.py checks these four against numbers the reference's
own Evaluator produced on the same lists (tests/golden).

`eval_tensors` is the device path: the top-k index tensor written by the scoring kernels is
scored against the test set by `eb_eval_topk_f64` without ever becoming Python tuples; same definitions, same
numbers (fp64; summation order differs, <= 1e-12 relative)."""
import math

import numpy as np

from .dataset import eval_csr_of

SUPPORTED = {"ndcg": "nDCG", "hr": "HR", "precision": "Precision", "recall": "Recall"}


class Evaluator:
    def __init__(self, data, params):
        self._data, self._params = data, params
        ev = data.config.evaluation
        self._k = getattr(ev, "cutoffs", [data.config.top_k])
        self._k = self._k if isinstance(self._k, list) else [self._k]
        if any(np.array(self._k) > data.config.top_k):
            raise Exception("Cutoff values must be smaller than recommendation list length (top_k)")
        self._metrics = []
        for m in ev.simple_metrics:
            if m.lower() not in SUPPORTED:
                raise Exception(f"metric {m} is not available in elliot_b200's evaluator "
                                f"(use the reference Evaluator through ProxyRecommender for the other 40)")
            self._metrics.append(SUPPORTED[m.lower()])
        self._sets = {"test": eval_csr_of(data, "test"), "val": eval_csr_of(data, "val")}

    def get_needed_recommendations(self):
        return self._data.config.top_k

    # recommendations: (val, test) pair of {public_user: [(public_item, score), ...]}
    def eval(self, recommendations):
        out = {}
        for k in self._k:
            res = {}
            for slot, which in ((0, "val"), (1, "test")):
                cs = self._sets[which]
                res[which] = None if cs is None else self._eval_dict(recommendations[slot], cs, k)
            if res["val"] is None:
                res["val"] = res["test"]
            if res["test"] is None:
                res["test"] = res["val"]
            out[k] = {"val_results": res["val"], "val_statistical_results": {},
                      "test_results": res["test"], "test_statistical_results": {}}
        return out

    # ---- device path ---------------------------------------------------------------------------
    def _device_set(self, which, k, device):
        """Item-sorted relevant-item CSR + per-user IDCG@k + discounts on `device` (cached)."""
        import torch
        key = (which, k, str(device))
        cache = self.__dict__.setdefault("_dev_sets", {})
        if key in cache:
            return cache[key]
        cs = self._sets[which]
        if cs is None:
            cache[key] = None
            return None
        indptr, idx, gain = cs
        n = len(indptr) - 1
        rows = np.repeat(np.arange(n), np.diff(indptr))
        disc = np.array([math.log(2) / math.log(r + 2) for r in range(k)])      # relevance.py:55
        by_gain = np.lexsort((-gain, rows))                                      # ideal ranking per user
        rank = np.arange(len(rows)) - indptr[rows]
        top = rank < k
        idcg = np.bincount(rows[top], weights=gain[by_gain][top] * disc[rank[top]], minlength=n)
        by_item = np.lexsort((idx, rows))                                        # lookup rows sorted by item id
        t = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a)).to(device=device, dtype=dt)
        cache[key] = (t(indptr, torch.int64), t(idx[by_item], torch.int32), t(gain[by_item], torch.float64),
                      t(idcg, torch.float64), t(disc, torch.float64))
        return cache[key]

    def eval_tensors(self, idx, users=None):
        """Same result structure as eval(), from a device (rows x top_k) int32 tensor of PRIVATE item ids
        (-1 = empty), row r = private user r (or users[r])."""
        from . import ops
        out = {}
        for k in self._k:
            res = {}
            for which in ("val", "test"):
                ds = self._device_set(which, k, idx.device)
                if ds is None:
                    res[which] = None
                    continue
                sums, _ = ops.eval_topk(idx, k, *ds, users=users)
                sums = sums.cpu().numpy()
                n = sums[0]
                vals = dict(zip(("nDCG", "HR", "Precision", "Recall"), (sums[1:] / n if n else np.zeros(4)).tolist()))
                res[which] = {m: vals[m] for m in self._metrics}
            if res["val"] is None:
                res["val"] = res["test"]
            if res["test"] is None:
                res["test"] = res["val"]
            out[k] = {"val_results": res["val"], "val_statistical_results": {},
                      "test_results": res["test"], "test_statistical_results": {}}
        return out

    def _eval_dict(self, recs, cs, k):
        pub_u, pub_i = self._data.public_users, self._data.public_items
        users = [u for u in recs if u in pub_u]
        idx = np.full((len(users), k), -1, np.int64)
        for r, u in enumerate(users):
            row = [pub_i.get(it, -1) for it, _ in recs[u][:k]]
            idx[r, :len(row)] = row
        return self.eval_arrays(np.array([pub_u[u] for u in users], np.int64), idx, cs, k)

    def eval_arrays(self, priv_users, idx, cs, k):
        """idx: (n, >=k) private item ids (−1 = empty slot), rows aligned with priv_users."""
        indptr, rel_idx, rel_gain = cs
        disc = np.array([math.log(2) / math.log(r + 2) for r in range(k)])      # relevance.py:55
        acc = {m: [] for m in self._metrics}
        for row, pu in zip(idx[:, :k], priv_users):
            lo, hi = indptr[pu], indptr[pu + 1]
            if hi == lo:                       # users without relevant test items are skipped
                continue
            gains = dict(zip(rel_idx[lo:hi].tolist(), rel_gain[lo:hi].tolist()))
            g = np.array([gains.get(int(it), 0.0) if it >= 0 else 0.0 for it in row])
            hits = g > 0
            if "nDCG" in acc:
                ideal = np.sort(rel_gain[lo:hi])[::-1][:k]
                idcg = float((ideal * disc[:len(ideal)]).sum())
                dcg = float((g * disc[:len(g)]).sum())
                acc["nDCG"].append(dcg / idcg if dcg > 0 else 0.0)
            if "HR" in acc:
                acc["HR"].append(1.0 if hits.any() else 0.0)
            # Precision/Recall sum the DISCOUNTED-relevance lookups? No: binary relevance
            # (precision.py / recall.py use relevance.binary_relevance.get_rel -> 1/0).
            if "Precision" in acc:
                acc["Precision"].append(hits.sum() / k)
            if "Recall" in acc:
                acc["Recall"].append(hits.sum() / (hi - lo))
        return {m: (float(np.mean(v)) if v else 0.0) for m, v in acc.items()}
