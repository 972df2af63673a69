"""MultiVAE step time at the C3 shape (ML-20M: 138493 x 26744, batch 512, 600/200)."""
import numpy as np, torch, time
from elliot_b200.recommender.multi_vae import VariationalAutoEncoder
dev = "cuda:0"
nu, ni, B = 138493, 26744, 512
g = torch.Generator(device=dev); g.manual_seed(0)
per = 144
cand = (torch.rand(nu, per, device=dev, generator=g) ** 2 * ni).to(torch.int32).clamp_(max=ni - 1)
cand, _ = torch.sort(cand, dim=1); keep = torch.ones_like(cand, dtype=torch.bool); keep[:, 1:] = cand[:, 1:] != cand[:, :-1]
indptr = torch.zeros(nu + 1, dtype=torch.int64, device=dev); indptr[1:] = torch.cumsum(keep.sum(1), 0); indices = cand[keep].contiguous()
m = VariationalAutoEncoder(ni, 600, 200, 1e-3, 0.0, 0.01, 42, indptr, indices, dev)
rows = torch.randperm(nu, device=dev, generator=g)[:B].to(torch.int32)
for _ in range(3): m.train_step(rows, 0.1)
torch.cuda.synchronize()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): loss = m.train_step(rows, 0.1)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
flop = 3 * 2 * B * ni * 600 + 3 * 2 * B * 600 * 400 + 3 * 2 * B * 200 * 600     # fwd+bwd dense (input layer is a gather)
print(f"MultiVAE C3 step: {ms:.3f} ms  {B/ms*1e3:.0f} users/s  {flop/ms/1e9:.1f} TFLOP/s dense-equivalent  loss {loss:.4f}")
