"""DataSet mirror: the fields of elliot/dataset/dataset.py:177-245 that the hot path consumes,
with the reference's id ordering (users in first-appearance order; private item index =
iteration order of a CPython set, dataset.py:201-202), plus the CSR tensors the kernels read.

Out of scope: side information, prefiltering, eval-time negative
sampling.  The dense `allunrated_mask` (dataset.py:245) is available lazily for code that wants
it; the kernels mask through the train CSR instead (a dense mask is 40 TB at C5 scale)."""
import logging
from types import SimpleNamespace

import numpy as np
import scipy.sparse as sp


class DataSet:
    def __init__(self, config, data_tuple, *args, **kwargs):
        self.logger = logging.getLogger("elliot_b200.DataSet")
        self.config = config
        self.args, self.kwargs = args, kwargs
        self.side_information = SimpleNamespace()
        self.train_dict = self._frame_to_dict(data_tuple[0])
        self.users = list(self.train_dict.keys())
        self.items = list({k for a in self.train_dict.values() for k in a.keys()})   # set order == reference
        self.num_users, self.num_items = len(self.users), len(self.items)
        self.transactions = sum(len(v) for v in self.train_dict.values())
        self.private_users = dict(enumerate(self.users))
        self.public_users = {v: k for k, v in self.private_users.items()}
        self.private_items = dict(enumerate(self.items))
        self.public_items = {v: k for k, v in self.private_items.items()}
        self.i_train_dict = {self.public_users[u]: {self.public_items[i]: r for i, r in its.items()}
                             for u, its in self.train_dict.items()}
        self.sp_i_train = self._csr(np.float32, ones=True)
        self.sp_i_train_ratings = self._csr(np.float32, ones=False)
        if len(data_tuple) == 2:
            self.test_dict = self._restrict(data_tuple[1])
        else:
            self.val_dict = self._restrict(data_tuple[1])
            self.test_dict = self._restrict(data_tuple[2])
        self._mask = None
        self._dev = {}

    # ---- construction helpers ------------------------------------------------------------
    @staticmethod
    def _frame_to_dict(df):
        """{user: {item: rating}} with users in first-appearance order and, per user, items in
        row order (what dataset.py:247-255 produces with its per-user filter loop)."""
        out = {}
        for u, i, r in zip(df["userId"].tolist(), df["itemId"].tolist(), df["rating"].tolist()):
            out.setdefault(u, {})[i] = r
        return out

    def _restrict(self, df):
        """test/val dict keyed by every TRAIN user (possibly empty), dataset.py:257-262."""
        raw = self._frame_to_dict(df)
        return {u: raw.get(u, {}) for u in self.users}

    def _csr(self, dtype, ones):
        rows, cols, vals = [], [], []
        for u, its in self.i_train_dict.items():
            for i, r in its.items():
                rows.append(u); cols.append(i); vals.append(1 if ones else r)
        return sp.csr_matrix((vals, (rows, cols)), dtype=dtype, shape=(self.num_users, self.num_items))

    # ---- reference accessors ---------------------------------------------------------------
    def get_test(self):
        return self.test_dict

    def get_validation(self):
        return getattr(self, "val_dict", None)

    @property
    def allunrated_mask(self):
        if self._mask is None:
            self._mask = np.where(self.sp_i_train.toarray() == 0, True, False)
        return self._mask

    # ---- kernel-side views (thin wrappers; the models call the free functions below, which also work on the
    # reference's own DataSet object when the plugin runs inside a real Elliot install) ----------------------
    def sampler_rows(self):
        return sampler_rows_of(self)

    def train_csr(self, device):
        return train_csr_of(self, device)

    def eval_csr(self, which="test"):
        return eval_csr_of(self, which)


# ---------------------------------------------------------------------------------------------------
# Kernel-side views of ANY object with the reference DataSet's fields (dataset.py:199-245): `i_train_dict`,
# `sp_i_train`, `users`, `public_items`, `test_dict` / `val_dict`, `config`.  They are free functions on purpose:
# inside a real Elliot install `self._data` is elliot.dataset.dataset.DataSet, which has none of the mirror's
# helper methods (the drop-in path must not depend on them).
# ---------------------------------------------------------------------------------------------------
def sampler_rows_of(data):
    """The reference sampler's `_ui_dict` (custom_sampler.py:21): per private user list(set(items)) — CPython set
    order, needed only where the reference's MT19937 stream is replayed (exact mode)."""
    return [list(set(data.i_train_dict[u])) for u in range(len(data.users))]


def train_csr_of(data, device, set_order=True):
    """(indptr int64, set-order indices int32 or None, sorted indices int32) on `device`, cached on the data object.
    The sorted CSR comes straight from `sp_i_train` (scipy, C speed — no per-row Python); the set-order copy is built
    row by row in Python because it IS CPython's set iteration order (exact mode only; set_order=False skips it)."""
    import torch
    cache = data.__dict__.setdefault("_eb200_dev", {})
    key = (str(device), bool(set_order))
    if key not in cache:
        other = cache.get((str(device), True))
        if other is not None:                                        # the full triple covers the partial request
            cache[key] = other
            return other
        m = data.sp_i_train.tocsr()
        if not m.has_sorted_indices:
            m = m.sorted_indices()
        indptr = m.indptr.astype(np.int64)
        srt = m.indices.astype(np.int32)
        flat = None
        if set_order:
            rows = sampler_rows_of(data)
            flat = np.fromiter((x for r in rows for x in r), dtype=np.int32, count=int(indptr[-1]))
            assert all(len(r) == indptr[u + 1] - indptr[u] for u, r in enumerate(rows)), "sp_i_train and i_train_dict disagree"
        cache[key] = tuple(None if a is None else torch.from_numpy(a).to(device) for a in (indptr, flat, srt))
    return cache[key]


def eval_csr_of(data, which="test"):
    """Host CSR (indptr, private item ids, gains) of the relevant items per private user (evaluator.py:117-147,
    relevance.py:80-82); None when the split does not exist."""
    d = data.test_dict if which == "test" else getattr(data, "val_dict", None)
    if d is None:
        return None
    thr = data.config.evaluation.relevance_threshold
    n_users = len(data.users)
    indptr = np.zeros(n_users + 1, np.int64)
    idx, gain = [], []
    for pu, u in enumerate(data.users):
        for it, score in d.get(u, {}).items():
            if score >= thr:
                idx.append(data.public_items.get(it, -1))            # test-only items can never be recommended
                gain.append(2 ** (score - thr + 1) - 1)              # relevance.py:80-82
        indptr[pu + 1] = len(idx)
    return indptr, np.array(idx, np.int64), np.array(gain, np.float64)
