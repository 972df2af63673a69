import sys, torch
from elliot_b200 import ops
dev = "cuda:0"
def run(nu, ni, d, bias):
    g = torch.Generator(device=dev); g.manual_seed(0)
    ld = ops.padded_dim(d)
    U = torch.zeros((nu, ld), device=dev); V = torch.zeros((ni, ld), device=dev)
    U[:, :d] = torch.randn(nu, d, device=dev, generator=g) * 0.1; V[:, :d] = torch.randn(ni, d, device=dev, generator=g) * 0.1
    b = torch.randn(ni, device=dev, generator=g) * 0.05 if bias else None
    for _ in range(2): i1, v1, st = ops.score_topk_tc(U, V, b, d, 10)
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): i1, v1, st = ops.score_topk_tc(U, V, b, d, 10)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print(f"nu={nu} ni={ni} d={d} bias={bias} kp={st['kp']}: {ms:.2f} ms {nu/ms/1e3:.3f} M users/s  alg {2.0*d*ni*nu/ms/1e9:.1f} TF/s  exec {2.0*st['kp']*ni*nu/ms/1e9:.1f} TF/s", flush=True)
for d, bias in [(128, False), (128, True), (64, False), (64, True), (200, True)]:
    run(148 * 128 * 2, 2_000_000, d, bias)
