"""Random 256-byte row gathers from a peer GPU's memory as a function of the REMOTE FOOTPRINT (torchrun, 2+ ranks).

At 8 GPUs the row gather of tools/peer_check.py drops from 615 GB/s (4 GPUs, 0.77 GB of peer memory touched at random) to 44 GB/s
(8 GPUs, 1.8 GB).  This probe holds the rank count fixed and grows the table instead, with the ids in random and in sorted order:
if the cliff follows the footprint and sorted ids recover the bandwidth, the limit is the reach of the address translation for
peer mappings, not the links."""
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
ld, n = 64, 1 << 22
res = []
for rows_per in (1_000_000, 2_000_000, 4_000_000, 8_000_000, 16_000_000):
    tab = PeerShardedTable(rows_per * world, ld, device=dev)
    tab.local.fill_(1.0)
    tab.barrier()
    g = torch.Generator(device=dev); g.manual_seed(rank)
    # remote rows only, so that the number is the link's
    own_lo = rank * tab.shard_rows
    ids = torch.randint(0, rows_per * (world - 1), (n,), device=dev, generator=g, dtype=torch.int64)
    ids = torch.where(ids >= own_lo, ids + tab.shard_rows, ids).to(torch.int32)
    dst = torch.empty((n, ld), device=dev)
    for order in ("random", "sorted"):
        use = ids if order == "random" else torch.sort(ids)[0].contiguous()
        for _ in range(2):
            ops.gather_rows_peer_f32(tab.ptrs, tab.shard_rows, ld, use, ld, out=dst)
        torch.cuda.synchronize(); dist.barrier()
        a, z = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(5):
            ops.gather_rows_peer_f32(tab.ptrs, tab.shard_rows, ld, use, ld, out=dst)
        z.record(); torch.cuda.synchronize()
        t = torch.tensor([a.elapsed_time(z) / 5], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        res.append({"remote_GB": rows_per * (world - 1) * 256 / 1e9, "order": order, "ms": t.item(), "GBps_in": n * 256 / t.item() / 1e6})
    tab.close(); del dst, ids
    torch.cuda.empty_cache()
if rank == 0:
    print(json.dumps({"world": world, "rows_gathered": n, "results": res}))
dist.destroy_process_group()
