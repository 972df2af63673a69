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

    # ---- kernel-side views -----------------------------------------------------------------
    def sampler_rows(self):
        """The reference sampler's `_ui_dict` (custom_sampler.py:21): per user list(set(items))."""
        return [list(set(self.i_train_dict[u])) for u in range(self.num_users)]

    def train_csr(self, device):
        """(indptr int64, set-order indices int32, sorted indices int32) on `device`."""
        import torch
        key = str(device)
        if key not in self._dev:
            rows = self.sampler_rows()
            indptr = np.zeros(self.num_users + 1, np.int64)
            indptr[1:] = np.cumsum([len(r) for r in rows])
            flat = np.fromiter((x for r in rows for x in r), dtype=np.int32, count=int(indptr[-1]))
            srt = np.fromiter((x for r in rows for x in sorted(r)), dtype=np.int32, count=int(indptr[-1]))
            self._dev[key] = tuple(torch.from_numpy(a).to(device) for a in (indptr, flat, srt))
        return self._dev[key]

    def eval_csr(self, which="test"):
        """Host CSR (indptr, private item ids, gains) of the relevant items per private user."""
        d = self.test_dict if which == "test" else getattr(self, "val_dict", None)
        if d is None:
            return None
        thr = self.config.evaluation.relevance_threshold
        indptr = np.zeros(self.num_users + 1, np.int64)
        idx, gain = [], []
        for pu, u in enumerate(self.users):
            for it, score in d.get(u, {}).items():
                if score >= thr:
                    idx.append(self.public_items.get(it, -1))       # test-only items can never be recommended
                    gain.append(2 ** (score - thr + 1) - 1)          # relevance.py:80-82
            indptr[pu + 1] = len(idx)
        return indptr, np.array(idx, np.int64), np.array(gain, np.float64)
