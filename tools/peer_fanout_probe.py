"""Random 256-byte row gathers over NVLink as a function of HOW MANY peers a GPU reads from at once (torchrun, 2+ ranks).

Modes: one      every id in the shard of rank+1
       all      ids uniform over all other shards (what a row-sharded table sees)
       grouped  the same ids, ordered by owner (each CTA wave talks to one peer at a time)
       sorted   the same ids fully sorted
The 8-GPU row gather of tools/peer_check.py runs at 44 GB/s where the 4-GPU one runs at 615 GB/s; tools/peer_tlb_probe.py shows that
the remote FOOTPRINT is not the reason (2 ranks: 640 GB/s from 0.25 to 4 GB)."""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from elliot_b200 import ops                          # noqa: E402
from elliot_b200.parallel import PeerShardedTable    # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
ld, n, rows_per = 64, 1 << 22, 1_000_000
tab = PeerShardedTable(rows_per * world, ld, device=dev)
tab.local.fill_(1.0)
tab.barrier()
g = torch.Generator(device=dev); g.manual_seed(rank)
own_lo = rank * tab.shard_rows
ids_all = torch.randint(0, rows_per * (world - 1), (n,), device=dev, generator=g, dtype=torch.int64)
ids_all = torch.where(ids_all >= own_lo, ids_all + tab.shard_rows, ids_all)
nxt = (rank + 1) % world
ids_one = torch.randint(0, rows_per, (n,), device=dev, generator=g, dtype=torch.int64) + nxt * tab.shard_rows
owner = ids_all // tab.shard_rows
# grouped: stable order by (owner - rank) mod world, so that rank r starts with peer r+1, then r+2, ... (no two ranks on one peer)
key = (owner - rank) % world
ids_grp = ids_all[torch.sort(key, stable=True)[1]]
modes = {"one": ids_one, "all": ids_all, "grouped": ids_grp, "sorted": torch.sort(ids_all)[0]}
dst = torch.empty((n, ld), device=dev)
res = {}
for name, ids in modes.items():
    use = ids.to(torch.int32).contiguous()
    for _ in range(2):
        ops.gather_rows_peer_f32(tab.ptrs, tab.shard_rows, ld, use, ld, out=dst)
    torch.cuda.synchronize(); dist.barrier()
    a, z = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5):
        ops.gather_rows_peer_f32(tab.ptrs, tab.shard_rows, ld, use, ld, out=dst)
    z.record(); torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(z) / 5], device=dev, dtype=torch.float64)
    tmin = t.clone()
    dist.all_reduce(t, op=dist.ReduceOp.MAX); dist.all_reduce(tmin, op=dist.ReduceOp.MIN)
    res[name] = {"ms_max": t.item(), "ms_min": tmin.item(), "GBps_in": n * 256 / t.item() / 1e6}
# one rank alone reads from all peers while the others idle
dist.barrier()
use = ids_all.to(torch.int32).contiguous()
if rank == 0:
    for _ in range(2):
        ops.gather_rows_peer_f32(tab.ptrs, tab.shard_rows, ld, use, ld, out=dst)
    torch.cuda.synchronize()
    a, z = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5):
        ops.gather_rows_peer_f32(tab.ptrs, tab.shard_rows, ld, use, ld, out=dst)
    z.record(); torch.cuda.synchronize()
    res["all_rank0_alone"] = {"ms": a.elapsed_time(z) / 5, "GBps_in": n * 256 / (a.elapsed_time(z) / 5) / 1e6}
dist.barrier()
if rank == 0:
    print(json.dumps({"world": world, "rows_gathered": n, "results": res}))
tab.close()
dist.destroy_process_group()
