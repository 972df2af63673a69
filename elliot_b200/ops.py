"""Torch-tensor wrappers over the C ABI.  Torch is plumbing only (device memory, streams);
all arithmetic happens in elliot_b200/csrc/*.cu.  Every op raises if its tensors are not on a
CUDA device — there is no CPU path.

Streams: every op launches on torch's CURRENT stream of the tensors' device.  The C ABI never allocates, so a few
wrappers keep a grow-only scratch buffer per device (scoring, evaluation, the native MultiVAE step, the MF2020
global-bias accumulator): those ops must not run concurrently on two streams of the same device from one process.
"""
import ctypes

import torch

from ._lib import check, lib

F32_STRIDES = (8, 16, 32, 64, 128, 256)


def padded_dim(d):
    """Row stride (in elements) the fp32 kernels need for `d` factors."""
    for s in F32_STRIDES:
        if d <= s:
            return s
    raise ValueError(f"factors={d} > 256 is not supported by the fp32 kernels")


def _ptr(t):
    return 0 if t is None else t.data_ptr()


def _stream(t):
    return torch.cuda.current_stream(t.device).cuda_stream


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("elliot_b200 ops need CUDA tensors (no CPU fallback exists)")


def _chk_idx(*ts):
    for t in ts:
        if t.dtype != torch.int32 or not t.is_contiguous():
            raise TypeError("index tensors must be contiguous int32")


def device_info():
    sm = ctypes.c_int(0); cc = ctypes.c_int(0)
    check(lib().eb_device_info(ctypes.byref(sm), ctypes.byref(cc)))
    return sm.value, cc.value


def bpr_step_f32(U, V, b, d, tu, ti, tj, lr, reg_u, reg_b, reg_pos, reg_neg, loss=None, racy=False):
    """Throughput-mode BPR step on materialised triples (BPRMF_model.py:87-117 semantics)."""
    _need_cuda(U, V, b, tu, ti, tj, loss)
    _chk_idx(tu, ti, tj)
    assert U.dtype == torch.float32 and V.dtype == torch.float32 and b.dtype == torch.float32
    assert U.stride(1) == 1 and V.stride(1) == 1 and U.stride(0) == V.stride(0)
    with torch.cuda.device(U.device):
        check(lib().eb_bpr_step_f32(_ptr(U), _ptr(V), _ptr(b), d, U.stride(0), _ptr(tu), _ptr(ti), _ptr(tj), tu.numel(),
                                    lr, reg_u, reg_b, reg_pos, reg_neg, _ptr(loss), 1 if racy else 0, _stream(U)))


def bloom_build(indptr, indices, n_users, words=32):
    """Per-user membership signatures for the fused sampler (32*words bits each, see eb_bloom_build): uint32 [n_users, words]
    (kept as int32 storage).  With them the sampler draws EXACTLY the same triples, just with a shorter load chain."""
    _need_cuda(indptr, indices)
    assert indptr.dtype == torch.int64 and indices.dtype == torch.int32
    out = torch.empty((n_users, words), dtype=torch.int32, device=indptr.device)
    with torch.cuda.device(indptr.device):
        check(lib().eb_bloom_build(_ptr(indptr), _ptr(indices), n_users, words, _ptr(out), _stream(indptr)))
    return out


def bpr_step_sampled_f32(U, V, b, d, n_users, n_items, indptr, indices, n, seed, first, lr, reg_u, reg_b, reg_pos,
                         reg_neg, loss=None, out=None, racy=False, reserve_sms=0, filter=None, _variant=0):
    """Fused sample+update step (custom_sampler.py:24-46 distribution, Philox stream).  filter: bloom_build() output.
    _variant (profiling): 16 forces the register-staged kernel, 32 the shared-memory-staged one (default: see use_stage)."""
    _need_cuda(U, V, b, indptr, indices, loss, filter)
    assert indptr.dtype == torch.int64 and indices.dtype == torch.int32
    ou = oi = oj = None
    if out is not None:
        ou, oi, oj = out
        _need_cuda(ou, oi, oj); _chk_idx(ou, oi, oj)
    with torch.cuda.device(U.device):
        check(lib().eb_bpr_step_sampled_filter_f32(_ptr(U), _ptr(V), _ptr(b), d, U.stride(0), n_users, n_items, _ptr(indptr),
                                                   _ptr(indices), _ptr(filter), 0 if filter is None else filter.shape[1], n, seed,
                                                   first, lr, reg_u, reg_b, reg_pos, reg_neg, _ptr(loss), _ptr(ou), _ptr(oi),
                                                   _ptr(oj), (1 if racy else 0) | ((int(reserve_sms) & 0xff) << 8) | int(_variant),
                                                   _stream(U)))


def bpr_sample_philox(n_users, n_items, indptr, indices, n, seed, first=0, filter=None):
    _need_cuda(indptr, indices, filter)
    dev = indptr.device
    u = torch.empty(n, dtype=torch.int32, device=dev); i = torch.empty_like(u); j = torch.empty_like(u)
    with torch.cuda.device(dev):
        check(lib().eb_bpr_sample_philox_filter(n_users, n_items, _ptr(indptr), _ptr(indices), _ptr(filter),
                                                0 if filter is None else filter.shape[1], n, seed, first, _ptr(u), _ptr(i), _ptr(j),
                                                _stream(indptr)))
    return u, i, j


def bpr_step_host_f32(U, V, b, d, tu_host, ti_host, tj_host, lr, reg_u, reg_b, reg_pos, reg_neg, staging, loss_dev,
                      loss_host, racy=False, sync=True, reserve_sms=0):
    """End-to-end step from HOST (pinned) int32 triples; returns after the loss is back on the host."""
    _need_cuda(U, V, b, staging, loss_dev)
    n = tu_host.numel()
    assert not tu_host.is_cuda and tu_host.dtype == torch.int32 and staging.numel() >= 3 * n
    with torch.cuda.device(U.device):
        check(lib().eb_bpr_step_host_f32(_ptr(U), _ptr(V), _ptr(b), d, U.stride(0), _ptr(tu_host), _ptr(ti_host),
                                         _ptr(tj_host), n, lr, reg_u, reg_b, reg_pos, reg_neg, _ptr(staging),
                                         _ptr(loss_dev), _ptr(loss_host), (1 if racy else 0) | (0 if sync else 2) | ((int(reserve_sms) & 0xff) << 8),
                                         _stream(U)))


def pack_bits(n_users, n_items):
    """(bits_u, bits_i) of the packed host-triple format."""
    bu, bi = max(1, (int(n_users) - 1).bit_length()), max(1, (int(n_items) - 1).bit_length())
    if bu + 2 * bi > 64:
        raise ValueError("ids too wide for the packed 64-bit triple format")
    return bu, bi


def pack_triples(tu, ti, tj, n_users, n_items):
    """int64 tensor (same device as the inputs): u | i << bits_u | j << (bits_u + bits_i) — the host-boundary format of
    bpr_step_host_packed_f32 (8 B/triple)."""
    bu, bi = pack_bits(n_users, n_items)
    return tu.to(torch.int64) | (ti.to(torch.int64) << bu) | (tj.to(torch.int64) << (bu + bi))


def bpr_step_host_packed_f32(U, V, b, d, packed_host, n_users, n_items, lr, reg_u, reg_b, reg_pos, reg_neg, staging, loss_dev,
                             loss_host, sync=True, reserve_sms=0):
    """End-to-end step from HOST (pinned) packed triples: one H2D copy of 8 B/triple, kernel, loss D2H."""
    _need_cuda(U, V, b, staging, loss_dev)
    n = packed_host.numel()
    assert not packed_host.is_cuda and packed_host.dtype == torch.int64 and staging.dtype == torch.int64 and staging.numel() >= n
    bu, bi = pack_bits(n_users, n_items)
    with torch.cuda.device(U.device):
        check(lib().eb_bpr_step_host_packed_f32(_ptr(U), _ptr(V), _ptr(b), d, U.stride(0), _ptr(packed_host), n, bu, bi,
                                                lr, reg_u, reg_b, reg_pos, reg_neg, _ptr(staging), _ptr(loss_dev), _ptr(loss_host),
                                                (0 if sync else 2) | ((int(reserve_sms) & 0xff) << 8), _stream(U)))


class _Workspace:
    """Grow-only device scratch buffer (the C ABI never allocates)."""

    def __init__(self):
        self.buf = None

    def get(self, nbytes, device):
        if self.buf is None or self.buf.numel() < nbytes or self.buf.device != device:
            self.buf = torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)
        return self.buf


_ws_exact, _ws_sampler, _ws_score = _Workspace(), _Workspace(), _Workspace()


def bpr_exact_f64(U, V, b, d, tu, ti, tj, lr, reg_u, reg_b, reg_pos, reg_neg, loss=None):
    """Exact-mode step: identical to applying the triples one by one in order (BPRMF.py:119-127)."""
    _need_cuda(U, V, b, tu, ti, tj, loss)
    _chk_idx(tu, ti, tj)
    assert U.dtype == torch.float64 and V.dtype == torch.float64 and b.dtype == torch.float64
    n = tu.numel()
    nu, ni = U.shape[0], V.shape[0]
    nbytes = lib().eb_bpr_exact_workspace_bytes(n, nu, ni)
    ws = _ws_exact.get(nbytes, U.device)
    with torch.cuda.device(U.device):
        check(lib().eb_bpr_exact_f64(_ptr(U), _ptr(V), _ptr(b), d, U.stride(0), nu, ni, _ptr(tu), _ptr(ti), _ptr(tj), n,
                                     lr, reg_u, reg_b, reg_pos, reg_neg, _ptr(loss), _ptr(ws), ws.numel(), _stream(U)))


class MtSampler:
    """Device replay of the reference sampler stream (custom_sampler.py:14-46).

    indptr/set_indices: the reference's `_ui_dict` in list(set(...)) order; sorted_indices: same
    CSR with rows sorted.  The MT19937 state persists across step() calls like the reference's
    global np.random state does across epochs.
    """

    def __init__(self, n_users, n_items, indptr, set_indices, sorted_indices, seed=42):
        _need_cuda(indptr, set_indices, sorted_indices)
        assert indptr.dtype == torch.int64 and set_indices.dtype == torch.int32 and sorted_indices.dtype == torch.int32
        self.n_users, self.n_items = n_users, n_items
        self.indptr, self.set_indices, self.sorted_indices = indptr, set_indices, sorted_indices
        self.state = torch.empty(625, dtype=torch.int32, device=indptr.device)
        with torch.cuda.device(indptr.device):
            check(lib().eb_mt_seed(_ptr(self.state), seed, _stream(indptr)))

    def raw(self, n):
        out = torch.empty(n, dtype=torch.int32, device=self.state.device)
        with torch.cuda.device(self.state.device):
            check(lib().eb_mt_raw(_ptr(self.state), _ptr(out), n, _stream(out)))
        return out

    def step(self, events):
        dev = self.state.device
        u = torch.empty(events, dtype=torch.int32, device=dev); i = torch.empty_like(u); j = torch.empty_like(u)
        ws = _ws_sampler.get(lib().eb_mt_sampler_workspace_bytes(events), dev)
        with torch.cuda.device(dev):
            check(lib().eb_mt_sampler_step(_ptr(self.state), self.n_users, self.n_items, _ptr(self.indptr),
                                           _ptr(self.set_indices), _ptr(self.sorted_indices), events, _ptr(u), _ptr(i),
                                           _ptr(j), _ptr(ws), ws.numel(), _stream(u)))
        return u, i, j


def score_topk(U, V, bias, d, k, mask_indptr=None, mask_indices=None, users=None, user_begin=0, n_sel=None):
    """Exact full-catalogue score + train mask + top-k (BPRMF_model.py:70-85 / BPRMF_batch_model.py:82-88)."""
    _need_cuda(U, V, bias, mask_indptr, mask_indices, users)
    if users is not None:
        _chk_idx(users)
        n_sel = users.numel()
    elif n_sel is None:
        n_sel = U.shape[0] - user_begin
    n_items = V.shape[0]
    dev = U.device
    idx = torch.empty((n_sel, k), dtype=torch.int32, device=dev)
    val = torch.empty((n_sel, k), dtype=U.dtype, device=dev)
    esz = U.element_size()
    ws = _ws_score.get(lib().eb_score_topk_workspace_bytes(n_sel, n_items, esz), dev)
    fn = lib().eb_score_topk_f32 if U.dtype == torch.float32 else lib().eb_score_topk_f64
    assert U.dtype in (torch.float32, torch.float64) and V.dtype == U.dtype
    with torch.cuda.device(dev):
        check(fn(_ptr(U), _ptr(V), _ptr(bias), n_items, d, U.stride(0), _ptr(mask_indptr), _ptr(mask_indices),
                 _ptr(users), user_begin, n_sel, k, _ptr(idx), _ptr(val), _ptr(ws), ws.numel(), _stream(U)))
    return idx, val


_ws_tc = _Workspace()


def score_topk_tc(U, V, bias, d, k, mask_indptr=None, mask_indices=None, user_begin=0, n_sel=None, dump=False, stats=True):
    """Tensor-core scoring + top-k (fp32 tables); same result as score_topk().  Returns
    (idx, val, stats) with stats = {"rechecked": users re-done by the exact kernel, "kp": padded K}
    (+ "dump": dense approximate scores when dump=True, tests only).  stats=False: nothing is read back, the call
    stays asynchronous (the models' path) and the dict is empty."""
    _need_cuda(U, V, bias, mask_indptr, mask_indices)
    assert U.dtype == torch.float32 and V.dtype == torch.float32 and U.stride(0) == V.stride(0)
    if n_sel is None:
        n_sel = U.shape[0] - user_begin
    n_items = V.shape[0]
    dev = U.device
    idx = torch.empty((n_sel, k), dtype=torch.int32, device=dev)
    val = torch.empty((n_sel, k), dtype=torch.float32, device=dev)
    dmp = torch.zeros((n_sel, n_items), dtype=torch.float32, device=dev) if dump else None
    ws = _ws_tc.get(lib().eb_score_topk_tc_workspace_bytes(n_sel, n_items, d), dev)
    st = (ctypes.c_int64 * 16)() if stats else None
    with torch.cuda.device(dev):
        check(lib().eb_score_topk_tc_f32(_ptr(U), _ptr(V), _ptr(bias), n_items, d, U.stride(0), _ptr(mask_indptr),
                                         _ptr(mask_indices), user_begin, n_sel, k, _ptr(idx), _ptr(val), _ptr(dmp),
                                         _ptr(ws), ws.numel(), ctypes.cast(st, ctypes.c_void_p) if stats else None, _stream(U)))
    out = {"rechecked": int(st[0]), "kp": int(st[1]), "prof": [int(x) for x in st[2:16]]} if stats else {}
    if dump:
        out["dump"] = dmp
    return idx, val, out


def bpr_batch_grad_f32(Gu, Gi, Bi, dGu, dGi, dBi, d, tu, ti, tj, l_w, l_b, loss=None):
    """Batch gradient of the BPRMF_batch loss (BPRMF_batch_model.py:57-75) into dense gradient tables."""
    _need_cuda(Gu, Gi, Bi, dGu, dGi, dBi, tu, ti, tj, loss)
    _chk_idx(tu, ti, tj)
    assert Gu.stride(0) == Gi.stride(0) == dGu.stride(0) == dGi.stride(0)
    with torch.cuda.device(Gu.device):
        check(lib().eb_bpr_batch_grad_f32(_ptr(Gu), _ptr(Gi), _ptr(Bi), _ptr(dGu), _ptr(dGi), _ptr(dBi), d, Gu.stride(0),
                                          _ptr(tu), _ptr(ti), _ptr(tj), tu.numel(), l_w, l_b, _ptr(loss), _stream(Gu)))


def adam_dense_f32(var, m, v, grad, lr, step, beta1=0.9, beta2=0.999, eps=1e-7):
    """Keras Adam over every element (TF 2.3 semantics for sparse gradients too); clears `grad`."""
    _need_cuda(var, m, v, grad)
    n = var.numel()
    assert var.is_contiguous() and m.numel() == n and v.numel() == n and grad.numel() == n and n % 4 == 0
    with torch.cuda.device(var.device):
        check(lib().eb_adam_dense_f32(_ptr(var), _ptr(m), _ptr(v), _ptr(grad), n, lr, beta1, beta2, eps, step,
                                      _stream(var)))


def table_delta_f32(cur, prev, delta):
    _need_cuda(cur, prev, delta)
    assert cur.is_contiguous() and prev.is_contiguous() and delta.is_contiguous() and cur.numel() % 4 == 0
    with torch.cuda.device(cur.device):
        check(lib().eb_table_delta_f32(_ptr(cur), _ptr(prev), _ptr(delta), cur.numel(), _stream(cur)))


def table_apply_delta_f32(cur, prev, delta_sum, scale=1.0):
    _need_cuda(cur, prev, delta_sum)
    with torch.cuda.device(cur.device):
        check(lib().eb_table_apply_delta_f32(_ptr(cur), _ptr(prev), _ptr(delta_sum), cur.numel(), scale, _stream(cur)))


_EXACT_GEMM = False


class exact_gemm:
    """Context manager (tests only): while active, `to_bf16` hands the fp32 tensors through unchanged and the dense layers run
    on the fp32 CUDA-core checking kernel (eb_gemm_f32_ref) instead of the bf16 tensor-core GEMM.  Models built inside the
    context can then be compared with their fp64 restatements to ~1e-5 — a check of the model WIRING that bf16 rounding would
    otherwise blur to 1e-2.  Slow; the product path never enables it; the native one-call MultiVAE step ignores it."""

    def __init__(self, on=True):
        self.on = on

    def __enter__(self):
        global _EXACT_GEMM
        self.prev, _EXACT_GEMM = _EXACT_GEMM, self.on
        return self

    def __exit__(self, *a):
        global _EXACT_GEMM
        _EXACT_GEMM = self.prev


def _gemm_ref(A, B, M, N, K, a_mn, b_mn, bias, alpha, act, out):
    assert A.dtype == torch.float32 and B.dtype == torch.float32
    if out is None:
        out = torch.empty((M, N), dtype=torch.float32, device=A.device)
    with torch.cuda.device(A.device):
        check(lib().eb_gemm_f32_ref(_ptr(A), A.stride(0), 1 if a_mn else 0, _ptr(B), B.stride(0), 1 if b_mn else 0, _ptr(out),
                                    out.stride(0), M, N, K, _ptr(bias), alpha, act, _stream(A)))
    return out


def to_bf16(src, transpose=False, out=None):
    """fp32 [R][C] -> bf16 [R][pad8(C)] or (transpose) [C][pad8(R)], zero padded, via eb_convert_bf16."""
    _need_cuda(src)
    if _EXACT_GEMM:                                       # checking mode: operands stay fp32
        return src.t().contiguous() if transpose else src
    assert src.dtype == torch.float32 and src.dim() == 2 and src.stride(1) == 1
    R, C = src.shape
    rows, cols = (C, R) if transpose else (R, C)
    ldd = (cols + 7) // 8 * 8
    if out is None:
        out = torch.empty((rows, ldd), dtype=torch.bfloat16, device=src.device)
    with torch.cuda.device(src.device):
        check(lib().eb_convert_bf16(_ptr(src), R, C, src.stride(0), _ptr(out), ldd, 1 if transpose else 0, _stream(src)))
    return out


def gemm_bf16_tn(A, B, M, N, K, bias=None, alpha=1.0, act=0, out=None, out_bf16=False):
    """C[M][N] fp32 = act(alpha * A[M][:K] @ B[N][:K]^T + bias) on the tensor cores (A, B bf16, K-major).
    out_bf16=True: returns (C, C_bf16) — the bf16 operand copy of C written by the epilogue itself (no conversion pass)."""
    _need_cuda(A, B, bias, out)
    if A.dtype == torch.float32:                          # exact_gemm checking mode
        C = _gemm_ref(A, B, M, N, K, False, False, bias, alpha, act, out)
        return (C, C) if out_bf16 else C
    if out_bf16:
        assert A.dtype == torch.bfloat16 and B.dtype == torch.bfloat16
        if out is None:
            out = torch.empty((M, N), dtype=torch.float32, device=A.device)
        ldb = (N + 7) // 8 * 8
        Cb = torch.empty((M, ldb), dtype=torch.bfloat16, device=A.device) if ldb == N else torch.zeros((M, ldb), dtype=torch.bfloat16, device=A.device)
        with torch.cuda.device(A.device):
            check(lib().eb_gemm_bf16_out(_ptr(A), A.stride(0), 0, _ptr(B), B.stride(0), 0, _ptr(out), out.stride(0), _ptr(Cb), ldb,
                                         M, N, K, _ptr(bias), alpha, act, _stream(A)))
        return out, Cb
    assert A.dtype == torch.bfloat16 and B.dtype == torch.bfloat16
    if out is None:
        out = torch.empty((M, N), dtype=torch.float32, device=A.device)
    with torch.cuda.device(A.device):
        check(lib().eb_gemm_bf16_tn(_ptr(A), A.stride(0), _ptr(B), B.stride(0), _ptr(out), out.stride(0), M, N, K, _ptr(bias),
                                    alpha, act, _stream(A)))
    return out


def gemm_bf16(A, B, M, N, K, a_rows_are_k=False, b_rows_are_k=False, bias=None, alpha=1.0, act=0, out=None):
    """C[M][N] fp32 = act(alpha * op(A) @ op(B)^T + bias): an operand flagged rows_are_k is a [K][M] (resp. [K][N]) row-major
    bf16 matrix read by the tensor cores as it lies (MN-major descriptors) — the backward GEMMs need no transposed copies."""
    _need_cuda(A, B, bias, out)
    if A.dtype == torch.float32:                          # exact_gemm checking mode
        return _gemm_ref(A, B, M, N, K, a_rows_are_k, b_rows_are_k, bias, alpha, act, out)
    assert A.dtype == torch.bfloat16 and B.dtype == torch.bfloat16
    if out is None:
        out = torch.empty((M, N), dtype=torch.float32, device=A.device)
    with torch.cuda.device(A.device):
        check(lib().eb_gemm_bf16(_ptr(A), A.stride(0), 1 if a_rows_are_k else 0, _ptr(B), B.stride(0), 1 if b_rows_are_k else 0,
                                 _ptr(out), out.stride(0), M, N, K, _ptr(bias), alpha, act, _stream(A)))
    return out


# ---------------------------------------------------------------- MultiVAE: whole step in one call (vae_step.cu)
class _VaeModelStruct(ctypes.Structure):
    _fields_ = ([("n_items", ctypes.c_int), ("H", ctypes.c_int), ("L", ctypes.c_int), ("reserved", ctypes.c_int)]
                + [(f"{pre}{k}", ctypes.c_void_p) for pre in ("", "g", "m", "v")
                   for k in ("W1", "b1", "W2", "b2", "W3", "b3", "W4", "b4")]
                + [(k, ctypes.c_void_p) for k in ("W2b", "W3b", "W4b", "W2t", "W3t", "W4t", "indptr", "indices")])


def vae_model_struct(n_items, H, L, P, G, M, V, bf16_copies, indptr, indices):
    """eb_vae_model over existing device tensors (which must outlive the struct and never be reallocated)."""
    _need_cuda(indptr, indices, *P.values(), *G.values(), *M.values(), *V.values(), *bf16_copies)
    st = _VaeModelStruct()
    st.n_items, st.H, st.L, st.reserved = n_items, H, L, 0
    for pre, d in (("", P), ("g", G), ("m", M), ("v", V)):
        for k in ("W1", "b1", "W2", "b2", "W3", "b3", "W4", "b4"):
            assert d[k].dtype == torch.float32 and d[k].is_contiguous()
            setattr(st, pre + k, d[k].data_ptr())
    for name, t in zip(("W2b", "W3b", "W4b"), bf16_copies):
        assert t.dtype == torch.bfloat16 and t.is_contiguous()
        setattr(st, name, t.data_ptr())
    st.W2t = st.W3t = st.W4t = None                      # unused since the backward GEMMs read the [out][in] copies directly
    assert indptr.dtype == torch.int64 and indices.dtype == torch.int32
    st.indptr, st.indices = indptr.data_ptr(), indices.data_ptr()
    st._keep = (P, G, M, V, bf16_copies, indptr, indices)
    return st


_vae_ws = {}


def vae_train_step(model, n_items, H, L, rows, drop_rate, noise_seed, drop_seed, step, anneal, lr, acc, phase=3):
    """eb_vae_train_step: phase bit 0 = forward+backward into the gradient buffers, bit 1 = Adam + operand refresh."""
    _need_cuda(acc, rows)
    dev = acc.device
    B = rows.numel() if rows is not None else 0
    ws = None
    if phase & 1:
        _chk_idx(rows)
        need = lib().eb_vae_step_workspace_bytes(n_items, H, L, B)
        ws = _vae_ws.get(dev)
        if ws is None or ws.numel() < need:
            ws = _vae_ws[dev] = torch.empty(need, dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        check(lib().eb_vae_train_step(ctypes.byref(model), _ptr(rows), B, drop_rate, noise_seed, drop_seed, step, anneal, lr,
                                      _ptr(acc), _ptr(ws), ws.numel() if ws is not None else 0, phase, _stream(acc)))


# ---------------------------------------------------------------- MultiVAE pieces (vae.cu)
def vae_embed_fwd(W1, b1, indptr, indices, rows, h1, drop_rate=0.0, seed=0):
    _need_cuda(W1, b1, indptr, indices, rows, h1)
    with torch.cuda.device(W1.device):
        check(lib().eb_vae_embed_fwd(_ptr(W1), _ptr(b1), W1.shape[1], _ptr(indptr), _ptr(indices), _ptr(rows), rows.numel(),
                                     _ptr(h1), h1.stride(0), drop_rate, seed, _stream(W1)))


def vae_embed_bwd(dW1, indptr, indices, rows, dpre1, drop_rate=0.0, seed=0):
    _need_cuda(dW1, indptr, indices, rows, dpre1)
    with torch.cuda.device(dW1.device):
        check(lib().eb_vae_embed_bwd(_ptr(dW1), dW1.shape[1], _ptr(indptr), _ptr(indices), _ptr(rows), rows.numel(), _ptr(dpre1),
                                     dpre1.stride(0), drop_rate, seed, _stream(dW1)))


def vae_reparam_fwd(ml, L, z, seed, step, kl_sum=None):
    _need_cuda(ml, z, kl_sum)
    with torch.cuda.device(ml.device):
        check(lib().eb_vae_reparam_fwd(_ptr(ml), ml.stride(0), ml.shape[0], L, _ptr(z), z.stride(0), seed, step, _ptr(kl_sum),
                                       _stream(ml)))


def vae_reparam_bwd(ml, L, dz, dml, seed, step, anneal):
    _need_cuda(ml, dz, dml)
    with torch.cuda.device(ml.device):
        check(lib().eb_vae_reparam_bwd(_ptr(ml), ml.stride(0), ml.shape[0], L, _ptr(dz), dz.stride(0), _ptr(dml), dml.stride(0),
                                       seed, step, anneal, _stream(ml)))


def vae_softmax(logits, indptr, indices, rows, nll_sum=None, lse_out=None, write_grad=True):
    _need_cuda(logits, indptr, indices, rows, nll_sum, lse_out)
    with torch.cuda.device(logits.device):
        check(lib().eb_vae_softmax(_ptr(logits), logits.stride(0), logits.shape[1], _ptr(indptr), _ptr(indices), _ptr(rows),
                                   logits.shape[0], _ptr(nll_sum), _ptr(lse_out), 1 if write_grad else 0, _stream(logits)))


def tanh_bwd(dout, out, dpre=None):
    _need_cuda(dout, out, dpre)
    assert dout.is_contiguous() and out.is_contiguous()
    if dpre is None:
        dpre = torch.empty_like(dout)
    with torch.cuda.device(dout.device):
        check(lib().eb_tanh_bwd(_ptr(dout), _ptr(out), _ptr(dpre), dout.numel(), _stream(dout)))
    return dpre


def colsum(src, out=None):
    _need_cuda(src, out)
    if out is None:
        out = torch.empty(src.shape[1], dtype=torch.float32, device=src.device)
    with torch.cuda.device(src.device):
        check(lib().eb_colsum(_ptr(src), src.shape[0], src.shape[1], src.stride(0), _ptr(out), _stream(src)))
    return out


def dense_topk(scores, k, mask_indptr=None, mask_indices=None, rows=None, shift=None):
    _need_cuda(scores, mask_indptr, mask_indices, rows, shift)
    n = scores.shape[0]
    idx = torch.empty((n, k), dtype=torch.int32, device=scores.device)
    val = torch.empty((n, k), dtype=torch.float32, device=scores.device)
    with torch.cuda.device(scores.device):
        check(lib().eb_dense_topk_f32(_ptr(scores), scores.stride(0), n, scores.shape[1], _ptr(mask_indptr), _ptr(mask_indices),
                                      _ptr(rows), _ptr(shift), k, _ptr(idx), _ptr(val), _stream(scores)))
    return idx, val


# ---------------------------------------------------------------- NeuMF pieces (neumf.cu)
def _call(name, dev_tensor, *args):
    with torch.cuda.device(dev_tensor.device):
        check(getattr(lib(), name)(*args, _stream(dev_tensor)))


def neumf_gather(Umf, Imf, Umlp, Imlp, f, u, it, x0, pm):
    _need_cuda(Umf, Imf, Umlp, Imlp, u, it, x0, pm)
    _call("eb_neumf_gather", Umf, _ptr(Umf), _ptr(Imf), _ptr(Umlp), _ptr(Imlp), f, Umf.stride(0), _ptr(u), _ptr(it), u.numel(),
          _ptr(x0), x0.stride(0), _ptr(pm), pm.stride(0))


def neumf_head(pm, h3, f, wp, bp, label=None, dpm=None, dh3=None, dwp=None, dbp=None, loss=None, prob=None, mean_over=None):
    """mean_over: number of samples the BinaryCrossentropy mean runs over (default: this call's batch)."""
    _need_cuda(pm, h3, wp, bp, label, dpm, dh3, dwp, dbp, loss, prob)
    _call("eb_neumf_head_norm", pm, _ptr(pm), pm.stride(0), _ptr(h3), h3.stride(0), f, _ptr(wp), _ptr(bp), _ptr(label), pm.shape[0],
          int(mean_over) if mean_over else max(int(pm.shape[0]), 1), _ptr(dpm), _ptr(dh3), _ptr(dwp), _ptr(dbp), _ptr(loss), _ptr(prob))


def relu_bwd(dout, out, copy_bf16=False):
    """dpre = dout * (out > 0); copy_bf16=True also returns dpre's bf16 operand copy, written in the same pass (row length % 8 == 0)."""
    _need_cuda(dout, out)
    dpre = torch.empty_like(dout)
    if copy_bf16 and not _EXACT_GEMM:
        assert dout.dim() == 2 and dout.shape[1] % 8 == 0 and dout.is_contiguous()
        cb = torch.empty(dout.shape, dtype=torch.bfloat16, device=dout.device)
        _call("eb_relu_bwd_copy", dout, _ptr(dout), _ptr(out), _ptr(dpre), dout.numel(), _ptr(cb))
        return dpre, cb
    _call("eb_relu_bwd", dout, _ptr(dout), _ptr(out), _ptr(dpre), dout.numel())
    return (dpre, dpre) if copy_bf16 else dpre


def neumf_scatter(Umf, Imf, f, u, it, dpm, dx0, dUmf, dImf, dUmlp, dImlp):
    _need_cuda(Umf, Imf, u, it, dpm, dx0, dUmf, dImf, dUmlp, dImlp)
    _call("eb_neumf_scatter", Umf, _ptr(Umf), _ptr(Imf), f, Umf.stride(0), _ptr(u), _ptr(it), u.numel(), _ptr(dpm), dpm.stride(0),
          _ptr(dx0), dx0.stride(0), _ptr(dUmf), _ptr(dImf), _ptr(dUmlp), _ptr(dImlp))


def neumf_sample(n_users, n_items, indptr, indices, m, seed):
    _need_cuda(indptr, indices)
    total = int(indices.numel()) * (1 + m)
    dev = indptr.device
    u = torch.empty(total, dtype=torch.int32, device=dev); i = torch.empty_like(u)
    y = torch.empty(total, dtype=torch.float32, device=dev)
    _call("eb_neumf_sample", indptr, n_users, n_items, _ptr(indptr), _ptr(indices), m, seed, total, _ptr(u), _ptr(i), _ptr(y))
    return u, i, y


def neumf_pair_h1(Au, Ai, b1, n_ub, n_items, h1, out):
    _need_cuda(Au, Ai, b1, out)
    if out.dtype == torch.float32:                        # exact_gemm checking mode: the first layer stays fp32
        _call("eb_neumf_pair_h1_f32", Au, _ptr(Au), Au.stride(0), _ptr(Ai), Ai.stride(0), _ptr(b1), n_ub, n_items, h1, _ptr(out), out.stride(0))
        return
    _call("eb_neumf_pair_h1", Au, _ptr(Au), Au.stride(0), _ptr(Ai), Ai.stride(0), _ptr(b1), n_ub, n_items, h1, _ptr(out), out.stride(0))


def neumf_pair_head(Umf, Imf, f, u0, n_ub, n_items, h3, wp, bp, prob):
    _need_cuda(Umf, Imf, h3, wp, bp, prob)
    _call("eb_neumf_pair_head", Umf, _ptr(Umf), _ptr(Imf), Umf.stride(0), f, u0, n_ub, n_items, _ptr(h3), h3.stride(0), _ptr(wp),
          _ptr(bp), _ptr(prob), prob.stride(0))


def table_apply_delta_late_f32(cur, prev, delta_sum, delta_local, scale=1.0):
    _need_cuda(cur, prev, delta_sum, delta_local)
    with torch.cuda.device(cur.device):
        check(lib().eb_table_apply_delta_late_f32(_ptr(cur), _ptr(prev), _ptr(delta_sum), _ptr(delta_local), cur.numel(),
                                                  float(scale), _stream(cur)))


def mf_pointwise_exact_f64(U, V, ub, ib, gb, d, su, si, sr, lr, reg, batch=100000, batch_loss=None):
    """MF2020 train_step over an ordered sample list, sequentially consistent fp64 (MF_model.py:80-112)."""
    _need_cuda(U, V, ub, ib, gb, su, si, sr, batch_loss)
    _chk_idx(su, si, sr)
    for t in (U, V, ub, ib, gb):
        assert t.dtype == torch.float64
    assert U.stride(1) == 1 and V.stride(1) == 1 and U.stride(0) == V.stride(0) and ub.is_contiguous() and ib.is_contiguous()
    n = su.numel()
    if batch_loss is not None:
        assert batch_loss.dtype == torch.float64 and batch_loss.numel() >= (n + batch - 1) // batch
    with torch.cuda.device(U.device):
        check(lib().eb_mf_pointwise_exact_f64(_ptr(U), _ptr(V), _ptr(ub), _ptr(ib), _ptr(gb), d, U.stride(0), _ptr(su), _ptr(si),
                                              _ptr(sr), n, lr, reg, batch, _ptr(batch_loss), _stream(U)))


_mf_gb_work = {}


def mf_pointwise_step_f32(U, V, ub, ib, gb, d, pos_u, pos_i, m, n_items, seed, epoch, lr, reg, loss=None, out=None,
                          first=0, count=None):
    """MF2020 throughput mode: positions [first, first+count) of the epoch's sample list (every positive + m uniform
    negatives, pseudo-randomly ordered) in one launch, fp32 Hogwild; count=None -> to the end of the epoch."""
    _need_cuda(U, V, ub, ib, gb, pos_u, pos_i, loss)
    _chk_idx(pos_u, pos_i)
    for t in (U, V, ub, ib, gb):
        assert t.dtype == torch.float32
    assert U.stride(1) == 1 and V.stride(1) == 1 and U.stride(0) == V.stride(0)
    ou = oi = orr = None
    if out is not None:
        ou, oi, orr = out
        _need_cuda(ou, oi, orr); _chk_idx(ou, oi, orr)
    n_epoch = pos_u.numel() * (1 + m)
    if count is None:
        count = n_epoch - first
    work = _mf_gb_work.get(U.device)
    if work is None:
        work = _mf_gb_work[U.device] = torch.zeros(4, dtype=torch.float64, device=U.device)
    with torch.cuda.device(U.device):
        check(lib().eb_mf_pointwise_step_f32(_ptr(U), _ptr(V), _ptr(ub), _ptr(ib), _ptr(gb), d, U.stride(0), _ptr(pos_u), _ptr(pos_i),
                                             pos_u.numel(), m, n_items, seed, epoch, first, count, lr, reg, _ptr(loss), _ptr(work),
                                             _ptr(ou), _ptr(oi), _ptr(orr), _stream(U)))


_eval_ws = None


def eval_topk(topk_idx, k, rel_indptr, rel_items, rel_gains, idcg, discount, users=None, per_user=False):
    """Accuracy metrics of a (rows x >=k) int32 top-k index tensor against an item-sorted relevant-item CSR
    (evaluator.py:117-147 semantics).  Returns (out, per_user): out = device double[5]
    {evaluated users, sum nDCG, sum HR, sum Precision, sum Recall}; per_user = (rows x 4) or None."""
    global _eval_ws
    _need_cuda(topk_idx, rel_indptr, rel_items, rel_gains, idcg, discount, users)
    assert topk_idx.dtype == torch.int32 and topk_idx.stride(1) == 1 and topk_idx.shape[1] >= k
    assert rel_indptr.dtype == torch.int64 and rel_items.dtype == torch.int32
    assert rel_gains.dtype == torch.float64 and idcg.dtype == torch.float64 and discount.dtype == torch.float64
    assert discount.numel() >= k
    n = topk_idx.shape[0]
    if users is not None:
        _chk_idx(users); assert users.numel() == n
    dev = topk_idx.device
    out = torch.empty(5, dtype=torch.float64, device=dev)
    pu = torch.empty(n, 4, dtype=torch.float64, device=dev) if per_user else None
    need = lib().eb_eval_topk_workspace_bytes(n, k)
    if _eval_ws is None or _eval_ws.numel() < need or _eval_ws.device != dev:
        _eval_ws = torch.empty(need, dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        check(lib().eb_eval_topk_f64(_ptr(topk_idx), n, topk_idx.stride(0), k, _ptr(users), _ptr(rel_indptr), _ptr(rel_items),
                                     _ptr(rel_gains), _ptr(idcg), _ptr(discount), _ptr(pu), _ptr(out), _ptr(_eval_ws),
                                     _eval_ws.numel(), _stream(topk_idx)))
    return out, pu


def partition_streams(device, reserve_sms, n_streams=1):
    """Streams bound to a green context that leaves >= reserve_sms SMs of `device` free (partition.cu).
    Returns (list of torch streams, SMs in the partition)."""
    arr = (ctypes.c_void_p * n_streams)()
    granted = ctypes.c_int(0)
    with torch.cuda.device(device):
        check(lib().eb_partition_streams_create(int(reserve_sms), n_streams, arr, ctypes.byref(granted)))
    return [torch.cuda.ExternalStream(int(arr[k]), device=device) for k in range(n_streams)], granted.value


# ---------------------------------------------------------------- row-sharded tables (sharded.cu)
def gather_rows_f32(table, ids, width=None, out=None):
    _need_cuda(table, ids, out); _chk_idx(ids)
    width = width or table.shape[1]
    if out is None:
        out = torch.empty((ids.numel(), table.stride(0)), dtype=torch.float32, device=table.device)
    _call("eb_gather_rows_f32", table, _ptr(table), table.stride(0), _ptr(ids), ids.numel(), width, _ptr(out), out.stride(0))
    return out


def scatter_add_rows_f32(table, ids, rows, width=None):
    _need_cuda(table, ids, rows); _chk_idx(ids)
    width = width or table.shape[1]
    _call("eb_scatter_add_rows_f32", table, _ptr(table), table.stride(0), _ptr(ids), ids.numel(), width, _ptr(rows), rows.stride(0))


def bpr_step_rows_f32(U, tu, Ri, Rj, bias_col, lr, reg_u, reg_b, reg_pos, reg_neg, loss=None):
    """BPR update against fetched item rows; returns (dRi, dRj) deltas to send back to the owners."""
    _need_cuda(U, tu, Ri, Rj, loss); _chk_idx(tu)
    dRi, dRj = torch.empty_like(Ri), torch.empty_like(Rj)
    _call("eb_bpr_step_rows_f32", U, _ptr(U), U.stride(0), _ptr(tu), _ptr(Ri), _ptr(Rj), Ri.stride(0), tu.numel(), bias_col,
          lr, reg_u, reg_b, reg_pos, reg_neg, _ptr(dRi), _ptr(dRj), _ptr(loss))
    return dRi, dRj


# ---------------------------------------------------------------- peer-addressed tables (csrc/peer.cu, bpr_train.cu PEER mode)
def _ptr_array(ptrs):
    """HOST array of device addresses (the `*_shards` arguments of the C ABI): ints, tensors, or a ready ctypes array."""
    if isinstance(ptrs, ctypes.Array):
        return ptrs, len(ptrs)
    vals = [p.data_ptr() if torch.is_tensor(p) else int(p) for p in ptrs]
    return (ctypes.c_void_p * len(vals))(*vals), len(vals)


def bpr_step_peer_f32(U, V_shards, b_shards, shard_rows, d, n_items, tu, ti, tj, lr, reg_u, reg_b, reg_pos, reg_neg, loss=None,
                      _variant=0):
    """BPR step on materialised triples, item table + biases row-sharded (shard s = rows [s*shard_rows, (s+1)*shard_rows)).
    _variant (profiling): 16 forces the register-staged kernel, 32 the shared-memory-staged one."""
    _need_cuda(U, tu, ti, tj, loss); _chk_idx(tu, ti, tj)
    va, n = _ptr_array(V_shards); ba, nb = _ptr_array(b_shards)
    assert n == nb and U.dtype == torch.float32 and U.stride(1) == 1
    with torch.cuda.device(U.device):
        check(lib().eb_bpr_step_peer_f32(_ptr(U), va, ba, n, shard_rows, d, U.stride(0), n_items, _ptr(tu), _ptr(ti), _ptr(tj),
                                         tu.numel(), lr, reg_u, reg_b, reg_pos, reg_neg, _ptr(loss), _variant, _stream(U)))


def bpr_step_sampled_peer_f32(U, V_shards, b_shards, shard_rows, d, n_users, n_items, indptr, indices, n, seed, first, lr, reg_u,
                              reg_b, reg_pos, reg_neg, loss=None, out=None, reserve_sms=0, filter=None, _no_item_updates=False,
                              _variant=0):
    """Fused sample+update step with the item table row-sharded over the GPUs of the box (loads / atomics over NVLink)."""
    _need_cuda(U, indptr, indices, loss)
    assert indptr.dtype == torch.int64 and indices.dtype == torch.int32 and U.dtype == torch.float32 and U.stride(1) == 1
    va, ns = _ptr_array(V_shards); ba, nb = _ptr_array(b_shards)
    assert ns == nb
    ou = oi = oj = None
    if out is not None:
        ou, oi, oj = out
        _need_cuda(ou, oi, oj); _chk_idx(ou, oi, oj)
    with torch.cuda.device(U.device):
        check(lib().eb_bpr_step_sampled_peer_f32(_ptr(U), va, ba, ns, shard_rows, d, U.stride(0), n_users, n_items, _ptr(indptr),
                                                 _ptr(indices), _ptr(filter), 0 if filter is None else filter.shape[1], n, seed, first, lr, reg_u, reg_b, reg_pos, reg_neg, _ptr(loss),
                                                 _ptr(ou), _ptr(oi), _ptr(oj), ((int(reserve_sms) & 0xff) << 8) | (4 if _no_item_updates else 0) | int(_variant),
                                                 _stream(U)))


def table_reconcile_peer_f32(slice_ptrs, prev_slice, scale, max_ctas=0):
    """One-kernel reconciliation of a replicated table's slice (see eb_table_reconcile_peer_f32)."""
    _need_cuda(prev_slice)
    pa, n = _ptr_array(slice_ptrs)
    with torch.cuda.device(prev_slice.device):
        check(lib().eb_table_reconcile_peer_f32(pa, n, _ptr(prev_slice), prev_slice.numel(), float(scale), int(max_ctas),
                                                _stream(prev_slice)))


def neumf_gather_peer(Umf, Umlp, I_shards, shard_rows, ldi, f, u, it, x0, pm):
    _need_cuda(Umf, Umlp, u, it, x0, pm)
    ia, n = _ptr_array(I_shards)
    _call("eb_neumf_gather_peer", Umf, _ptr(Umf), _ptr(Umlp), Umf.stride(0), ia, n, shard_rows, ldi, f, _ptr(u), _ptr(it), u.numel(),
          _ptr(x0), x0.stride(0), _ptr(pm), pm.stride(0))


def neumf_scatter_peer(Umf, I_shards, GI_shards, shard_rows, ldi, f, u, it, dpm, dx0, dUmf, dUmlp):
    _need_cuda(Umf, u, it, dpm, dx0, dUmf, dUmlp)
    ia, n = _ptr_array(I_shards); ga, ng = _ptr_array(GI_shards)
    assert n == ng
    _call("eb_neumf_scatter_peer", Umf, _ptr(Umf), Umf.stride(0), ia, ga, n, shard_rows, ldi, f, _ptr(u), _ptr(it), u.numel(),
          _ptr(dpm), dpm.stride(0), _ptr(dx0), dx0.stride(0), _ptr(dUmf), _ptr(dUmlp))


def group_by_owner(arrays, key1, key2, shard_rows, rank, world):
    """Reorder up to three 32-bit arrays that travel together so that elements whose row (arrays[key1], and arrays[key2] if
    key2 >= 0) lives on the same owner are adjacent, starting with rank+1's rows (see eb_group_by_owner_i32: with more than four
    GPUs a peer kernel fed in this order moves rows 12x faster).  float32 payloads are carried bit for bit.  Returns new tensors."""
    arrays = list(arrays) + [None] * (3 - len(arrays))
    _need_cuda(*[x for x in arrays if x is not None])
    views = [None if x is None else x.view(torch.int32) for x in arrays]
    n = views[0].numel()
    outs = [None if v is None else torch.empty_like(v) for v in views]
    work = torch.empty(128, dtype=torch.int32, device=views[0].device)
    _call("eb_group_by_owner_i32", views[0], _ptr(views[0]), _ptr(views[1]), _ptr(views[2]), key1, key2, n, shard_rows, rank, world,
          _ptr(outs[0]), _ptr(outs[1]), _ptr(outs[2]), _ptr(work))
    return [None if o is None else o.view(x.dtype) for o, x in zip(outs, arrays)]


def gather_rows_peer_f32(shards, shard_rows, ld, ids, width, out=None):
    _need_cuda(ids, out); _chk_idx(ids)
    sa, n = _ptr_array(shards)
    if out is None:
        out = torch.empty((ids.numel(), width), dtype=torch.float32, device=ids.device)
    _call("eb_gather_rows_peer_f32", ids, sa, n, shard_rows, ld, _ptr(ids), ids.numel(), width, _ptr(out), out.stride(0))
    return out


# ---------------------------------------------------------------- GMF (csrc/gmf.cu)
def pointwise_sample_philox(n_users, n_items, indptr, indices, n, seed, first=0, filter=None):
    """(u, item, label) samples of pointwise_pos_neg_sampler.py:24-48 (Philox stream)."""
    _need_cuda(indptr, indices, filter)
    dev = indptr.device
    u = torch.empty(n, dtype=torch.int32, device=dev); i = torch.empty_like(u); y = torch.empty(n, dtype=torch.float32, device=dev)
    _call("eb_pointwise_sample_philox", indptr, n_users, n_items, _ptr(indptr), _ptr(indices), _ptr(filter),
          0 if filter is None else filter.shape[1], n, seed, first, _ptr(u), _ptr(i), _ptr(y))
    return u, i, y


def gmf_step_grads(U, I, h, f, u, it, label, dU, dI, dh, loss=None, mean_over=None):
    _need_cuda(U, I, h, u, it, label, dU, dI, dh, loss); _chk_idx(u, it)
    assert U.stride(0) == I.stride(0) == dU.stride(0) == dI.stride(0)
    _call("eb_gmf_step_grads", U, _ptr(U), _ptr(I), U.stride(0), f, _ptr(h), _ptr(u), _ptr(it), _ptr(label), u.numel(),
          int(mean_over) if mean_over else max(int(u.numel()), 1), _ptr(dU), _ptr(dI), _ptr(dh), _ptr(loss))


def gmf_scale_rows(src, h, f, out=None):
    _need_cuda(src, h, out)
    if out is None:
        out = torch.zeros_like(src)
    _call("eb_gmf_scale_rows", src, _ptr(src), src.stride(0), src.shape[0], f, _ptr(h), _ptr(out), out.stride(0))
    return out


def sigmoid_(x):
    _need_cuda(x)
    assert x.is_contiguous() and x.dtype == torch.float32
    _call("eb_sigmoid_inplace", x, _ptr(x), x.numel())
    return x
