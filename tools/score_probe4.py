import os, sys, torch
from elliot_b200 import ops
dev = "cuda:0"
nu, ni, d, k = 148 * 128, 100_000, int(sys.argv[1]) if len(sys.argv) > 1 else 64, 10
g = torch.Generator(device=dev); g.manual_seed(0)
U = torch.randn(nu, d, device=dev, generator=g) * 0.1; V = torch.randn(ni, d, device=dev, generator=g) * 0.1
for _ in range(2): ops.score_topk_tc(U, V, None, d, k)
torch.cuda.synchronize()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3): ops.score_topk_tc(U, V, None, d, k)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 3
BN = 256 if d <= 128 else (128 if d <= 192 else 64)
print(f"EB_TC_DEBUG={os.environ.get('EB_TC_DEBUG','0')} d={d}: {ms:.3f} ms  {ms*1e3/((ni+BN-1)//BN):.2f} us/tile")
