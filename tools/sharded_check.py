"""torchrun --nproc-per-node N tools/sharded_check.py : row-sharded item table over N GPUs (NCCL all-to-all).
(1) correctness: a conflict-free batch gives the same tables as the single-table kernel; (2) throughput at a
C4-like per-GPU shape (users local, items sharded)."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from elliot_b200 import ops
from elliot_b200.parallel import ShardedTable, shard_range, sharded_bpr_step
os.environ.setdefault("NCCL_DEBUG", "WARN")
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
hp = (0.05, 0.0025, 0.0, 0.0025, 0.00025)
# ---- (1) correctness on a small conflict-free batch: every rank uses disjoint users and disjoint items
ni, d, ld, nu_loc = 4096, 64, 64, 512
g = torch.Generator(device=dev); g.manual_seed(5)                     # same V on every rank
V = torch.randn(ni, ld, device=dev, generator=g) * 0.1
gu = torch.Generator(device=dev); gu.manual_seed(100 + rank)
U = torch.randn(nu_loc, ld, device=dev, generator=gu) * 0.1
lo, hi = shard_range(ni, rank, world)
items = ShardedTable(ni, V[lo:hi].clone())
n = 256
tu = torch.arange(n, dtype=torch.int32, device=dev)
perm = torch.randperm(ni, device=dev, generator=g)                   # same permutation everywhere
ti = perm[rank * 2 * n: rank * 2 * n + n].to(torch.int32); tj = perm[rank * 2 * n + n: (rank + 1) * 2 * n].to(torch.int32)
U_ref, V_ref = U.clone(), V.clone(); b_ref = torch.zeros(ni, device=dev)
ops.bpr_step_f32(U_ref, V_ref, b_ref, d, tu, ti, tj, *hp)          # single-table kernel on a private full copy
sharded_bpr_step(U, items, tu, ti, tj, hp)
torch.cuda.synchronize(); dist.barrier()
err_u = (U - U_ref).abs().max().item()
# rows of my shard touched by ANY rank must match that rank's reference: gather references
touched = torch.zeros(ni, dtype=torch.bool, device=dev); touched[ti.long()] = True; touched[tj.long()] = True
mine = V_ref.clone(); mine[~touched] = 0; dist.all_reduce(mine)      # sum of the touched rows (disjoint across ranks)
cnt = touched.float(); dist.all_reduce(cnt)
expect = torch.where(cnt[:, None] > 0, mine, V)
err_v = (items.local - expect[lo:hi]).abs().max().item()
ok = err_u < 1e-6 and err_v < 1e-6
# ---- (2) throughput: 1M local users x (1M x world) sharded items, d=64, 1M triples per step per rank
nu_loc, ni_tot, B = 1_000_000, 1_000_000 * world, 1 << 20
U = torch.randn(nu_loc, ld, device=dev, generator=gu) * 0.1
lo, hi = shard_range(ni_tot, rank, world)
items = ShardedTable(ni_tot, torch.randn(hi - lo, ld, device=dev, generator=gu) * 0.1)
tu = torch.randint(0, nu_loc, (B,), device=dev, generator=gu, dtype=torch.int32)
ti = torch.randint(0, ni_tot, (B,), device=dev, generator=gu, dtype=torch.int32); tj = torch.randint(0, ni_tot, (B,), device=dev, generator=gu, dtype=torch.int32)
for _ in range(3): sharded_bpr_step(U, items, tu, ti, tj, hp)
torch.cuda.synchronize(); dist.barrier()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True); e0.record()
K = 10
for _ in range(K): sharded_bpr_step(U, items, tu, ti, tj, hp)
e1.record(); torch.cuda.synchronize()
t = torch.tensor([e0.elapsed_time(e1) / K], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    print(json.dumps({"world": world, "conflict_free_max_err_user": err_u, "conflict_free_max_err_item": err_v, "ok": ok,
                      "sharded_step_ms": t.item(), "triples_per_s_all_ranks": B * world / (t.item() * 1e-3),
                      "shape": f"{nu_loc} local users, {ni_tot} items sharded over {world} GPUs, d=64, {B} triples/step/rank",
                      "nvlink_bytes_per_triple_per_direction": 2 * (4 + 256 + 256) * (world - 1) / world}))
dist.destroy_process_group()
