"""Ad-hoc device probe: times the BPR kernels at C2 scale (not the bench contract)."""
import sys, time
import numpy as np, torch
from elliot_b200 import ops

dev = "cuda:0"
nu, ni, d = 1_000_000, 100_000, 64
per_user = 100
torch.manual_seed(0)
U = (torch.randn(nu, d, device=dev) * 0.1); V = (torch.randn(ni, d, device=dev) * 0.1); b = torch.zeros(ni, device=dev)
# synthetic CSR: per user `per_user` candidates, squared-uniform popularity skew, dedup by sort
cand = (torch.rand(nu, per_user, device=dev) ** 2 * ni).to(torch.int32).clamp_(max=ni - 1)
cand, _ = torch.sort(cand, dim=1)
keep = torch.ones_like(cand, dtype=torch.bool); keep[:, 1:] = cand[:, 1:] != cand[:, :-1]
lens = keep.sum(1)
indptr = torch.zeros(nu + 1, dtype=torch.int64, device=dev); indptr[1:] = torch.cumsum(lens, 0)
indices = cand[keep].contiguous()
print("nnz", indices.numel(), "mean len", lens.float().mean().item())
hp = (0.05, 0.0025, 0.0, 0.0025, 0.00025)
B = 1 << 22
loss = torch.zeros(1, dtype=torch.float64, device=dev)
out = [torch.empty(B, dtype=torch.int32, device=dev) for _ in range(3)]

def timeit(fn, iters=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters

step = [0]
def sampled(racy=False, emit=False):
    step[0] += 1
    ops.bpr_step_sampled_f32(U, V, b, d, nu, ni, indptr, indices, B, 42, step[0] * B, *hp, loss=loss, out=out if emit else None, racy=racy)

for name, fn in [("sampled atomic", lambda: sampled()), ("sampled racy", lambda: sampled(True)), ("sampled atomic emit", lambda: sampled(False, True))]:
    ms = timeit(fn)
    print(f"{name:24s} {ms:8.3f} ms  {B/ms/1e6:8.3f} G triples/s  alg {B*1552/ms/1e6:8.1f} GB/s")
u, i, j = ops.bpr_sample_philox(nu, ni, indptr, indices, B, 7)
ms = timeit(lambda: ops.bpr_sample_philox(nu, ni, indptr, indices, B, 7))
print(f"{'philox sample only':24s} {ms:8.3f} ms  {B/ms/1e6:8.3f} G triples/s")
for name, racy in [("materialised atomic", False), ("materialised racy", True)]:
    ms = timeit(lambda: ops.bpr_step_f32(U, V, b, d, u, i, j, *hp, loss=loss, racy=racy))
    print(f"{name:24s} {ms:8.3f} ms  {B/ms/1e6:8.3f} G triples/s  alg {B*1564/ms/1e6:8.1f} GB/s")
print("finite:", torch.isfinite(U).all().item(), torch.isfinite(V).all().item(), "loss", loss.item())
# exact mode throughput on ML-1M-like shape
nu2, ni2, T = 6040, 3706, 800_000
U2 = torch.randn(nu2, d, device=dev, dtype=torch.float64) * 0.1; V2 = torch.randn(ni2, d, device=dev, dtype=torch.float64) * 0.1
b2 = torch.zeros(ni2, device=dev, dtype=torch.float64)
pop = (torch.rand(T, device=dev) ** 2 * ni2).to(torch.int32).clamp_(max=ni2 - 1)
tu = torch.randint(0, nu2, (T,), device=dev, dtype=torch.int32); tj = torch.randint(0, ni2, (T,), device=dev, dtype=torch.int32)
tj = torch.where(tj == pop, (tj + 1) % ni2, tj).to(torch.int32)
ms = timeit(lambda: ops.bpr_exact_f64(U2, V2, b2, d, tu, pop, tj, *hp), iters=3, warm=1)
print(f"exact f64 ML-1M-like epoch {ms:8.3f} ms  {T/ms/1e3:8.3f} M triples/s")
