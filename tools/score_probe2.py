import sys, torch
from elliot_b200 import ops
dev = "cuda:0"
def run(nu, ni, d, k, per_user, use_bias, use_mask, bias_scale=0.05):
    g = torch.Generator(device=dev); g.manual_seed(0)
    ld = ops.padded_dim(d)
    U = torch.zeros((nu, ld), device=dev); V = torch.zeros((ni, ld), device=dev)
    U[:, :d] = torch.randn(nu, d, device=dev, generator=g) * 0.1; V[:, :d] = torch.randn(ni, d, device=dev, generator=g) * 0.1
    b = torch.randn(ni, device=dev, generator=g) * bias_scale if use_bias else None
    indptr = indices = None
    if use_mask:
        cand = (torch.rand(nu, per_user, device=dev, generator=g) ** 2 * ni).to(torch.int32).clamp_(max=ni - 1)
        cand, _ = torch.sort(cand, dim=1); keep = torch.ones_like(cand, dtype=torch.bool); keep[:, 1:] = cand[:, 1:] != cand[:, :-1]
        indptr = torch.zeros(nu + 1, dtype=torch.int64, device=dev); indptr[1:] = torch.cumsum(keep.sum(1), 0); indices = cand[keep].contiguous()
    for _ in range(2): i1, v1, st = ops.score_topk_tc(U, V, b, d, k, indptr, indices)
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); reps = 3
    for _ in range(reps): i1, v1, st = ops.score_topk_tc(U, V, b, d, k, indptr, indices)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    kp = st["kp"]; BN = 256 if kp <= 128 else 128
    tiles = ((nu + 127) // 128 + 147) // 148 * ((ni + BN - 1) // BN)
    print(f"nu={nu} ni={ni} d={d} bias={use_bias}({bias_scale}) mask={use_mask}: {ms:.2f} ms  {nu/ms*1e3/1e6:.3f} M users/s  {2.0*kp*ni*nu/ms/1e9:.1f} TFLOP/s  {ms*1e3/tiles:.2f} us/tile rechecked {st['rechecked']}", flush=True)
nu = 148 * 128
for ub, um in [(False, False), (True, False), (False, True), (True, True)]:
    run(nu, 100_000, 64, 10, 100, ub, um)
run(nu, 100_000, 64, 10, 100, True, False, bias_scale=0.001)
run(nu, 100_000, 128, 10, 100, False, False)
run(nu, 100_000, 256, 10, 100, False, False)
