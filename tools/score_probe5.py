import os, sys, torch
os.environ["EB_TC_PROF"] = "1"
from elliot_b200 import ops
dev = "cuda:0"
nu, ni, d, k = 148 * 128, int(sys.argv[2]) if len(sys.argv) > 2 else 100_000, int(sys.argv[1]) if len(sys.argv) > 1 else 64, 10
g = torch.Generator(device=dev); g.manual_seed(0)
U = torch.randn(nu, d, device=dev, generator=g) * 0.1; V = torch.randn(ni, d, device=dev, generator=g) * 0.1
for _ in range(2): i, v, st = ops.score_topk_tc(U, V, None, d, k)
names = ["wait_acc", "tmem_ld", "scan", "compact", "final+rerank", "n_compact_rows", "n_slow_chunks", "n_groups"]
BN = 256 if d <= 128 else (128 if d <= 192 else 64)
tiles = (ni + BN - 1) // BN
print(f"d={d} ni={ni} tiles={tiles} chunks={tiles*BN//32}")
for n, x in zip(names, st["prof"]): print(f"  {n:16s} {x:12d}  per tile {x/tiles:10.1f}")
