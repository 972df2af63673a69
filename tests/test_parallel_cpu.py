"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: shard partition, replicated-table
reconciliation protocol, user-sharded top-k gather."""
import os
import tempfile

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from elliot_b200.parallel import (GradAllReduce, OverlappedTableSync, ReplicatedTableSync, ShardedTable, gather_topk, owner_of,
                                  shard_range)


def test_shard_range_partitions():
    for n in [0, 1, 7, 8, 1000003]:
        for w in [1, 2, 3, 8]:
            rs = [shard_range(n, r, w) for r in range(w)]
            assert rs[0][0] == 0 and rs[-1][1] == n
            assert all(rs[i][1] == rs[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in rs]
            assert max(sizes) - min(sizes) <= 1
    for n, w in [(10, 3), (7, 2), (1000, 8)]:
        for row in range(n):
            r = owner_of(row, n, w)
            lo, hi = shard_range(n, r, w)
            assert lo <= row < hi


def _worker(rank, world, init_file, out_dir):
    dist.init_process_group("gloo", init_method=f"file://{init_file}", rank=rank, world_size=world)
    torch.manual_seed(0)
    V = torch.randn(16, 8); b = torch.randn(16)
    V0, b0 = V.clone(), b.clone()
    sync = ReplicatedTableSync([V, b], delta_fn=lambda c, p, d: d.copy_(c - p),
                               apply_fn=lambda c, p, s, k: (p.add_(s * k), c.copy_(p)), reduce="sum")
    for step in range(3):
        # each rank touches different rows with different amounts (its local Hogwild step)
        V[rank * 4 + step] += (rank + 1) * 0.5
        V[15] += 0.25                                            # a row BOTH ranks touch
        b[rank] -= 1.0
        sync.sync()
    want_V = V0.clone(); want_b = b0.clone()
    for step in range(3):
        for r in range(world):
            want_V[r * 4 + step] += (r + 1) * 0.5
            want_V[15] += 0.25
            want_b[r] -= 1.0
    ok = torch.allclose(V, want_V, atol=1e-6) and torch.allclose(b, want_b, atol=1e-6) and torch.equal(V, sync.prev[0])
    # user-sharded top-k gather keeps user order
    n_users = 11
    lo, hi = shard_range(n_users, rank, world)
    idx = torch.arange(lo, hi, dtype=torch.int32).unsqueeze(1).repeat(1, 3)
    val = idx.float() * 0.5
    gi, gv = gather_topk(idx, val, n_users)
    ok = ok and torch.equal(gi[:, 0], torch.arange(n_users, dtype=torch.int32)) and torch.equal(gv, gi.float() * 0.5)
    # averaged reduction through one flat buffer: replicas agree and hold the mean of the rank updates
    flat = torch.zeros(12)
    A, Bv = flat[:8].view(2, 4), flat[8:]
    msync = ReplicatedTableSync([A, Bv], delta_fn=lambda c, p, d: d.copy_(c - p),
                                apply_fn=lambda c, p, s, k: (p.add_(s * k), c.copy_(p)), reduce="mean", flat=flat)
    A += float(rank + 1); Bv -= 2.0 * rank
    msync.sync()
    ok = ok and torch.allclose(A, torch.full((2, 4), sum(r + 1 for r in range(world)) / world)) \
        and torch.allclose(Bv, torch.full((4,), -2.0 * sum(range(world)) / world))
    # overlapped (one-step-late) variant: after flush() the replicas agree and nothing is lost or doubled
    W = torch.zeros(8, 4)
    osync = OverlappedTableSync([W], delta_fn=lambda c, p, d: d.copy_(c - p),
                                late_fn=lambda c, p, s, l, k: (c.add_(s * k - l), p.add_(s * k)), reduce="sum")
    for step in range(4):
        W[rank + step] += 1.0 + rank
        W[7] += 0.5
        osync.sync()
    osync.flush()
    want_W = torch.zeros(8, 4)
    for step in range(4):
        for r in range(world):
            want_W[r + step] += 1.0 + r
            want_W[7] += 0.5
    ok = ok and torch.allclose(W, want_W, atol=1e-6) and torch.allclose(osync.prev[0], want_W, atol=1e-6)
    # overlapped + averaged + flat: replicas end identical and equal the running mean of the rank updates
    oflat = torch.zeros(12)
    OA, Ob = oflat[:8].view(2, 4), oflat[8:]
    omsync = OverlappedTableSync([OA, Ob], delta_fn=lambda c, p, d: d.copy_(c - p),
                                 late_fn=lambda c, p, s, l, k: (c.add_(s * k - l), p.add_(s * k)), reduce="mean", flat=oflat)
    for step in range(3):
        OA += float(rank + 1); Ob -= 2.0 * rank
        omsync.sync()
    omsync.flush()
    ok = ok and torch.allclose(OA, torch.full((2, 4), 3.0 * sum(r + 1 for r in range(world)) / world), atol=1e-6) \
        and torch.allclose(Ob, torch.full((4,), -6.0 * sum(range(world)) / world), atol=1e-6) \
        and torch.allclose(omsync.prev[0], oflat, atol=1e-6)
    # row-sharded table: fetch arbitrary global rows (with duplicates), push deltas back to the owners
    n_rows = 13
    full = torch.arange(n_rows * 4, dtype=torch.float32).reshape(n_rows, 4)
    lo, hi = shard_range(n_rows, rank, world)
    st = ShardedTable(n_rows, full[lo:hi].clone(), gather_fn=lambda t, i: t[i.long()].clone(),
                      scatter_fn=lambda t, i, r: t.index_add_(0, i.long(), r))
    ids = torch.tensor([12, 0, 5, 5, 7, rank, 11 - rank], dtype=torch.int32)
    got = st.fetch(ids)
    ok = ok and torch.equal(got, full[ids.long()])
    st.push(torch.ones(len(ids), 4) * (rank + 1))
    dist.barrier()
    want = full.clone()
    for r in range(world):
        rid = torch.tensor([12, 0, 5, 5, 7, r, 11 - r])
        want.index_add_(0, rid, torch.ones(len(rid), 4) * (r + 1))
    ok = ok and torch.equal(st.local, want[lo:hi])
    # push into a separate target (the dense gradient shard of an Adam-trained table): the table itself stays put
    before = st.local.clone()
    gshard = torch.zeros_like(st.local)
    st.fetch(ids)
    st.push(torch.ones(len(ids), 4) * (rank + 1), target=gshard)
    dist.barrier()
    ok = ok and torch.equal(st.local, before) and torch.equal(gshard, (want - full)[lo:hi])
    # data-parallel dense grads: one flat buffer averaged, loss accumulators summed
    gflat = torch.arange(6, dtype=torch.float32) * (rank + 1)
    acc = torch.tensor([1.0 + rank, 10.0], dtype=torch.float64)
    w = GradAllReduce(gflat, extra=acc).sync()
    ok = ok and w == world and torch.allclose(gflat, torch.arange(6, dtype=torch.float32) * sum(r + 1 for r in range(world)) / world) \
        and torch.equal(acc, torch.tensor([sum(1.0 + r for r in range(world)), 10.0 * world], dtype=torch.float64))
    # unequal and EMPTY slices (tail batch of an epoch): mean-over-own-rows gradients combined into the
    # global-batch mean; a rank without rows contributes zeros and still takes part
    X = torch.arange(12, dtype=torch.float32).reshape(3, 4) + 1.0           # 3 rows over 2 ranks: slices of 2 and 1
    for n_rows in (3, 1):                                                     # n_rows = 1: rank 1 is empty
        lo_r, hi_r = shard_range(n_rows, rank, world)
        mine = X[lo_r:hi_r]
        g = mine.mean(0).clone() if hi_r > lo_r else torch.zeros(4)
        GradAllReduce(g).sync_weighted(hi_r - lo_r, n_rows)
        ok = ok and torch.allclose(g, X[:n_rows].mean(0), atol=1e-6)
    torch.save({"ok": ok, "V": V}, os.path.join(out_dir, f"r{rank}.pt"))
    dist.destroy_process_group()


def test_replicated_table_sync_and_gather_world2():
    world = 2
    with tempfile.TemporaryDirectory() as d:
        init = os.path.join(d, "init")
        mp.spawn(_worker, args=(world, init, d), nprocs=world, join=True)
        res = [torch.load(os.path.join(d, f"r{r}.pt")) for r in range(world)]
    assert all(r["ok"] for r in res)
    assert torch.equal(res[0]["V"], res[1]["V"])                  # replicas identical after every sync
