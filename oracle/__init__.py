"""oracle — CPU restatement of the reference hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package, and only as the checker.
The product package ``elliot_b200`` never imports it (tests/test_capi_symbols.py::test_product_never_imports_oracle
enforces that).

Parity pinning: the reference ships no tests or golden vectors, so
``oracle/gen_golden.py`` runs the reference's own ``custom_sampler.py`` /
``BPRMF_model.py`` (by file path, in the build container) and commits the outputs
under ``tests/golden/``; ``tests/test_oracle_golden.py`` checks this restatement
against them; the MF2020 restatement (``mf2020_update_seq`` / ``mf2020_epoch_samples``) is pinned the same way
against ``tests/golden/mf2020_*.npz`` (reference MF2020 sampler + model run).  The TF-based variants (BPRMF_batch, NeuMF, MultiVAE) cannot be run
(tensorflow==2.3.2 is absent): their restatements in ``oracle/tf_models.py`` are
"parity unpinned" and say so.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle.so")
_SRC = os.path.join(_HERE, "bprmf_oracle.c")


def build(force=False):
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(_SRC):
        subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-ffp-contract=off", "-o", _SO, _SRC, "-lm"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = ctypes.CDLL(build())
        L.orc_rng_size.restype = ctypes.c_size_t
        L.orc_randint.restype = ctypes.c_int64
        L.orc_randint.argtypes = [ctypes.c_void_p, ctypes.c_int64]
        L.orc_sampler_step.restype = ctypes.c_int64
        L.orc_bpr_loss.restype = ctypes.c_double
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


class Rng:
    """numpy legacy RandomState restated (MT19937 + masked randint + legacy gauss)."""

    def __init__(self, seed):
        self._buf = ctypes.create_string_buffer(lib().orc_rng_size())
        lib().orc_seed(self._buf, ctypes.c_uint32(seed))

    def randint(self, n):
        return int(lib().orc_randint(self._buf, ctypes.c_int64(n)))

    def raw(self, n):
        out = np.empty(n, dtype=np.uint32)
        lib().orc_raw_fill(self._buf, _p(out), ctypes.c_int64(n))
        return out

    def normal(self, loc, scale, size):
        out = np.empty(int(np.prod(size)), dtype=np.float64)
        lib().orc_normal_fill(self._buf, ctypes.c_double(loc), ctypes.c_double(scale), _p(out),
                              ctypes.c_int64(out.size))
        return out.reshape(size)


def mf_init(seed, n_users, n_items, d):
    """MFModel.initialize (BPRMF_model.py:38-56)."""
    U = np.empty((n_users, d)); V = np.empty((n_items, d)); b = np.empty(n_items)
    lib().orc_mf_init(ctypes.c_uint32(seed), ctypes.c_int64(n_users), ctypes.c_int64(n_items), ctypes.c_int(d),
                      _p(U), _p(V), _p(b))
    return U, V, b


def sampler_step(rng, n_users, n_items, indptr, indices, events):
    """Sampler.step (custom_sampler.py:24-46) -> (u, i, j) int32 arrays, raw draws consumed."""
    indptr = np.ascontiguousarray(indptr, dtype=np.int64)
    indices = np.ascontiguousarray(indices, dtype=np.int32)
    u = np.empty(events, np.int32); i = np.empty(events, np.int32); j = np.empty(events, np.int32)
    draws = lib().orc_sampler_step(rng._buf, ctypes.c_int64(n_users), ctypes.c_int64(n_items), _p(indptr), _p(indices),
                                   ctypes.c_int64(events), _p(u), _p(i), _p(j))
    if draws < 0:
        raise RuntimeError("a user owns every item: the reference sampler never terminates")
    return u, i, j, int(draws)


def bpr_update_seq(U, V, b, u, i, j, lr=0.05, reg_u=0.0025, reg_b=0.0, reg_pos=0.0025, reg_neg=0.00025):
    """MFModel.update_factors looped (BPRMF_model.py:87-117); in place on float64 C arrays."""
    assert U.dtype == np.float64 and U.flags.c_contiguous and V.flags.c_contiguous
    u = np.ascontiguousarray(u, np.int32); i = np.ascontiguousarray(i, np.int32); j = np.ascontiguousarray(j, np.int32)
    lib().orc_bpr_update_seq(_p(U), _p(V), _p(b), ctypes.c_int(U.shape[1]), ctypes.c_double(lr),
                             ctypes.c_double(reg_u), ctypes.c_double(reg_b), ctypes.c_double(reg_pos),
                             ctypes.c_double(reg_neg), _p(u), _p(i), _p(j), ctypes.c_int64(len(u)))


def bpr_loss(U, V, b, u, i, j):
    u = np.ascontiguousarray(u, np.int32); i = np.ascontiguousarray(i, np.int32); j = np.ascontiguousarray(j, np.int32)
    return float(lib().orc_bpr_loss(_p(U), _p(V), _p(b), ctypes.c_int(U.shape[1]), _p(u), _p(i), _p(j),
                                    ctypes.c_int64(len(u))))


def user_topk(U, V, b, mask_indptr, mask_indices, users, k):
    """MFModel.get_user_predictions (BPRMF_model.py:70-85) for a list of private user ids."""
    U = np.ascontiguousarray(U, np.float64); V = np.ascontiguousarray(V, np.float64)
    b = None if b is None else np.ascontiguousarray(b, np.float64)
    users = np.ascontiguousarray(users, np.int32)
    idx = np.empty((len(users), k), np.int32); val = np.empty((len(users), k), np.float64)
    if mask_indptr is not None:
        mask_indptr = np.ascontiguousarray(mask_indptr, np.int64)
        mask_indices = np.ascontiguousarray(mask_indices, np.int32)
    lib().orc_user_topk(_p(U), _p(V), None if b is None else _p(b), ctypes.c_int64(V.shape[0]), ctypes.c_int(U.shape[1]),
                        None if mask_indptr is None else _p(mask_indptr),
                        None if mask_indptr is None else _p(mask_indices),
                        _p(users), ctypes.c_int64(len(users)), ctypes.c_int(k), _p(idx), _p(val))
    return idx, val


def mf2020_update_seq(U, V, ub, ib, gb, su, si, sr, lr, reg, batch=100000):
    """MF2020 MFModel.train_step over a whole epoch's sample list, in order (MF_model.py:80-112).
    Updates U, V, ub, ib in place; returns (new global bias, per-batch loss sums)."""
    n = len(su)
    g = ctypes.c_double(gb)
    bl = np.zeros((n + batch - 1) // batch if n else 0, dtype=np.float64)
    for a in (U, V, ub, ib):
        assert a.dtype == np.float64 and a.flags.c_contiguous
    su, si, sr = (np.ascontiguousarray(x, dtype=np.int32) for x in (su, si, sr))
    lib().orc_mf2020_update_seq(_p(U), _p(V), _p(ub), _p(ib), ctypes.byref(g), ctypes.c_int(U.shape[1]),
                                ctypes.c_double(lr), ctypes.c_double(reg), _p(su), _p(si), _p(sr), ctypes.c_int64(n),
                                ctypes.c_int64(batch), _p(bl))
    return g.value, bl


def mf2020_epoch_samples(np_rs, py_rng, pos_u, pos_i, n_items, m):
    """custom_sampler_rendle.Sampler.step (MF2020/custom_sampler_rendle.py:29-85): every positive (u, i, 1) in
    sp_i_train.nonzero() order followed by m uniform items (u, randint(n_items), 0) — NOT rejected against the
    train set — then one random.sample permutation of the whole list.  np_rs: numpy legacy RandomState carrying
    the process's global stream position; py_rng: random.Random at the reference's position."""
    n = len(pos_u)
    mat = np.empty((n * (1 + m), 3), dtype=np.int32)
    mat[::1 + m, 0] = pos_u; mat[::1 + m, 1] = pos_i; mat[::1 + m, 2] = 1
    if m:
        neg = np_rs.randint(n_items, size=n * m).reshape(n, m)
        for q in range(m):
            mat[1 + q::1 + m, 0] = pos_u; mat[1 + q::1 + m, 1] = neg[:, q]; mat[1 + q::1 + m, 2] = 0
    perm = py_rng.sample(range(len(mat)), len(mat))
    return mat[perm]
