"""Interpreter-bound NumPy restatement of the reference's BPRMF hot loop.

TEST/BENCH INFRASTRUCTURE ONLY (see oracle/__init__.py).  Purpose: a CPU baseline whose cost
profile matches the reference's (one Python-level call and a handful of small float64 NumPy
operations per triple, legacy global np.random stream) — the C port in bprmf_oracle.c is two
orders of magnitude faster than anything the reference itself can do on a host core.

Restates  elliot/dataset/samplers/custom_sampler.py:24-46  (draw order: user, positive slot,
negative with redraw-on-hit)  and  elliot/recommender/latent_factor_models/BPRMF/
BPRMF_model.py:66-117  (fp64 sequential SGD incl. the updated-user-row aliasing).  Checked
against tests/golden in tests/test_oracle_golden.py::test_numpy_port_matches_golden.
"""
import numpy as np


def triple_stream(rows, n_items, events):
    """Yield `events` (u, i, j) triples from the GLOBAL legacy numpy stream (caller seeds it).

    rows[u] is the user's train-item list in the reference's list(set(...)) order."""
    draw = np.random.randint
    n_users = len(rows)
    sizes = [len(r) for r in rows]
    produced = 0
    while produced < events:
        u = draw(n_users)
        mine = rows[u]
        if sizes[u] >= n_items:
            raise RuntimeError("user owns every item: the reference never terminates here")
        pos = mine[draw(sizes[u])]
        neg = draw(n_items)
        while neg in mine:          # list scan, like the reference
            neg = draw(n_items)
        produced += 1
        yield u, pos, neg


class SequentialBPR:
    def __init__(self, n_users, n_items, factors, lr, reg_u, reg_b, reg_pos, reg_neg, seed):
        np.random.seed(seed)
        self.hp = (lr, reg_u, reg_b, reg_pos, reg_neg)
        self.bias = np.zeros(n_items)
        self.P = np.random.normal(0, 0.1, (n_users, factors))   # users first, then items (BPRMF_model.py:53-56)
        self.Q = np.random.normal(0, 0.1, (n_items, factors))

    def score(self, u, it):
        return self.bias[it] + self.P[u] @ self.Q[it]

    def sgd(self, u, i, j):
        lr, reg_u, reg_b, reg_pos, reg_neg = self.hp
        pu, qi, qj = self.P[u], self.Q[i], self.Q[j]           # views: pu aliases the table row
        bi, bj = self.bias[i], self.bias[j]
        z = 1 / (1 + np.exp(self.score(u, i) - self.score(u, j)))
        self.bias[i] = bi + lr * (z - reg_b * bi)
        self.bias[j] = bj + lr * (-z - reg_b * bj)
        self.P[u] = pu + lr * ((qi - qj) * z - reg_u * pu)      # pu now shows the NEW row
        self.Q[i] = qi + lr * (pu * z - reg_pos * qi)
        self.Q[j] = qj + lr * (-pu * z - reg_neg * qj)

    def recommend(self, u, seen, k):
        s = self.bias + self.P[u] @ self.Q.T
        s[seen] = -np.inf
        pairs = [(it, val) for it, val in enumerate(s)]          # per-item Python loop, as in the reference
        idx = np.array([p[0] for p in pairs]); val = np.array([p[1] for p in pairs])
        kk = min(k, len(val))
        cand = np.argpartition(val, -kk)[-kk:]
        order = val[cand].argsort()[::-1]
        return [(int(idx[cand][o]), float(val[cand][o])) for o in order]
