import torch
from elliot_b200 import ops
dev = "cuda:0"
nu, ni, d, k = 148 * 128, 25600, 64, 10
g = torch.Generator(device=dev); g.manual_seed(0)
U = torch.randn(nu, d, device=dev, generator=g) * 0.1; V = torch.randn(ni, d, device=dev, generator=g) * 0.1
for _ in range(2): ops.score_topk_tc(U, V, None, d, k)
torch.cuda.synchronize()
