"""Multi-GPU plumbing for the hot path — one process per GPU, torch.distributed.

Training: user rows are sharded (rank r owns a contiguous block of users and samples triples
only for them, so user rows never move); the item table and item biases are REPLICATED and
reconciled once per step by one all-reduce of the per-rank deltas (the path's only exchange
step).  Scoring: users are sharded, V replicated, no collective.  The reference has no
distributed code at all; exact mode is single-GPU by definition.
"""
import torch
import torch.distributed as dist


def shard_range(n, rank, world):
    """Contiguous block partition of n rows: sizes differ by at most one, concatenation covers [0, n)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def owner_of(row, n, world):
    base, rem = divmod(n, world)
    cut = rem * (base + 1)
    return row // (base + 1) if row < cut else rem + (row - cut) // max(base, 1)


class ReplicatedTableSync:
    """Keeps replicated tables consistent across ranks:  T <- T_prev + scale * sum_r (T_r - T_prev).

    reduce="mean" (default, scale = 1/world): local-SGD style averaging — stable however many ranks hit the same
    rows with stale reads; reduce="sum": every rank's updates applied in full (what one GPU would have applied for
    the same triples), which overshoots when many ranks update the same hot rows from the same stale state
    (observed: non-finite tables at 4 ranks x 4M triples on a 100K-item table with lr 0.05).
    `flat` (optional): one contiguous tensor of which all `tables` are views -> ONE delta kernel, ONE all-reduce,
    ONE apply kernel per sync.
    delta_fn / apply_fn default to the CUDA kernels (ops.table_delta_f32 / table_apply_delta_f32);
    the CPU (gloo) tests inject torch equivalents to exercise the protocol without a GPU.
    """

    def __init__(self, tables, group=None, delta_fn=None, apply_fn=None, reduce="mean", flat=None):
        from . import ops
        assert reduce in ("mean", "sum")
        self.tables = [flat] if flat is not None else list(tables)
        self.prev = [t.clone() for t in self.tables]
        self.delta = [torch.empty_like(t) for t in self.tables]
        self.group, self.reduce = group, reduce
        self._delta = delta_fn or ops.table_delta_f32
        self._apply = apply_fn or ops.table_apply_delta_f32

    def reset(self):
        """Re-baseline on the tables' current contents (they must already agree across ranks)."""
        for t, p in zip(self.tables, self.prev):
            p.copy_(t)

    def sync(self):
        world = dist.get_world_size(self.group) if dist.is_initialized() else 1
        scale = 1.0 / world if self.reduce == "mean" else 1.0
        for t, p, d in zip(self.tables, self.prev, self.delta):
            self._delta(t, p, d)
            if world > 1:
                dist.all_reduce(d, group=self.group)
            self._apply(t, p, d, scale)


class GradAllReduce:
    """Data-parallel dense weights (MultiVAE encoder/decoder, NeuMF MLP — SURVEY.md §8e): every rank holds the same
    weights, computes gradients on its slice of the batch, and the gradients are AVERAGED with one all-reduce over
    one flat buffer before the (identical) optimizer step.  `extra` (optional): small tensors summed along the way
    (loss accumulators)."""

    def __init__(self, flat, group=None, extra=None):
        self.flat, self.group, self.extra = flat, group, extra

    def sync(self):
        if not dist.is_initialized():
            return 1
        world = dist.get_world_size(self.group)
        if world == 1:
            return 1
        if dist.get_backend(self.group) == "nccl":
            dist.all_reduce(self.flat, op=dist.ReduceOp.AVG, group=self.group)      # the division happens inside NCCL
        else:
            dist.all_reduce(self.flat, group=self.group)
            self.flat.div_(world)
        if self.extra is not None:
            dist.all_reduce(self.extra, group=self.group)
        return world

    def sync_weighted(self, local_rows, global_rows):
        """The same for slices of UNEQUAL size (the tail batch of an epoch; a rank may even hold no row at all and then
        contributes the zero gradient): every rank's gradient is a mean over ITS rows, the global-batch gradient is
        sum_r (B_r / B) g_r = mean_r (world B_r / B) g_r, so each rank pre-scales and the usual average follows.
        Every rank must call it, whatever its slice."""
        world = dist.get_world_size(self.group) if dist.is_initialized() else 1
        if world > 1 and int(global_rows) != int(local_rows) * world:
            self.flat.mul_(world * float(local_rows) / float(global_rows))
        return self.sync()


class OverlappedTableSync:
    """Same reconciliation, one step late, so the all-reduce of step k overlaps the compute of step k+1
    (Hogwild tolerates the extra staleness; nothing is lost or double counted):

        after step k  :  local_k = T - T_prev        (compute stream);  all_reduce(copy of local_k) on the comm stream
        after step k+1:  wait for that all-reduce;  T += scale*sum_k - local_k;  T_prev += scale*sum_k;  start round k+1

    The collective only overlaps if the compute kernel leaves SMs free for it: launch the persistent training
    grid with `reserve_sms` (flags bits 8..15 of eb_bpr_step_sampled_f32) >= the collective's CTA count.
    reduce / flat: as in ReplicatedTableSync.  delta_fn / late_fn default to the CUDA kernels; CPU tests inject
    torch equivalents.
    """

    def __init__(self, tables, group=None, delta_fn=None, late_fn=None, stream=None, reduce="mean", flat=None):
        from . import ops
        assert reduce in ("mean", "sum")
        self.tables = [flat] if flat is not None else list(tables)
        self.prev = [t.clone() for t in self.tables]
        self.local = [torch.zeros_like(t) for t in self.tables]
        self.sum = [torch.zeros_like(t) for t in self.tables]
        self.group, self.reduce = group, reduce
        self._delta = delta_fn or ops.table_delta_f32
        self._late = late_fn or ops.table_apply_delta_late_f32
        self.cuda = self.tables[0].is_cuda
        self.comm = stream if stream is not None else (torch.cuda.Stream(device=self.tables[0].device) if self.cuda else None)
        self.pending = False
        self.ready = torch.cuda.Event() if self.cuda else None
        self.done = torch.cuda.Event() if self.cuda else None

    def _world(self):
        return dist.get_world_size(self.group) if dist.is_initialized() else 1

    def _finish(self):
        if not self.pending:
            return
        if self.cuda:
            torch.cuda.current_stream(self.tables[0].device).wait_event(self.done)
        scale = 1.0 / self._world() if self.reduce == "mean" else 1.0
        for t, p, s, l in zip(self.tables, self.prev, self.sum, self.local):
            self._late(t, p, s, l, scale)
        self.pending = False

    def sync(self):
        """Call once after every local step."""
        self._finish()                                       # apply round k-1 (other ranks' updates arrive one step late)
        for t, p, l, s in zip(self.tables, self.prev, self.local, self.sum):
            self._delta(t, p, l)
            s.copy_(l)
        world = self._world()
        if self.cuda:
            self.ready.record()
            with torch.cuda.stream(self.comm):
                self.comm.wait_event(self.ready)
                if world > 1:
                    for s in self.sum:
                        dist.all_reduce(s, group=self.group)
                self.done.record()
        elif world > 1:
            for s in self.sum:
                dist.all_reduce(s, group=self.group)
        self.pending = True

    def flush(self):
        """Apply the outstanding round (end of training / before scoring)."""
        self._finish()

    def reset(self):
        """Flush, then re-baseline on the tables' current contents (they must already agree across ranks)."""
        self._finish()
        for t, p in zip(self.tables, self.prev):
            p.copy_(t)


class ShardedTable:
    """A table whose rows are block-partitioned over the ranks (`shard_range`).  `fetch(ids)` returns copies of
    arbitrary global rows (ids all-to-all -> owners gather -> rows all-to-all), `push(deltas)` sends per-row
    deltas for the ids of the last fetch back to their owners, which add them into their shard.

    gather_fn(local_table, local_ids) -> rows and scatter_fn(local_table, local_ids, rows) default to the CUDA
    kernels (ops.gather_rows_f32 / ops.scatter_add_rows_f32); CPU tests inject torch equivalents.
    Index bookkeeping (owner lookup, sort by owner, split sizes) uses torch ops: plumbing, no arithmetic on rows.
    """

    def __init__(self, n_rows, local, group=None, gather_fn=None, scatter_fn=None):
        from . import ops
        self.n_rows, self.local, self.group = n_rows, local, group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.lo, self.hi = shard_range(n_rows, self.rank, self.world)
        assert local.shape[0] == self.hi - self.lo
        self.bounds = torch.tensor([shard_range(n_rows, r, self.world)[1] for r in range(self.world)],
                                   dtype=torch.int64, device=local.device)
        self._gather = gather_fn or (lambda t, i: ops.gather_rows_f32(t, i))
        self._scatter = scatter_fn or (lambda t, i, r: ops.scatter_add_rows_f32(t, i, r))
        self._plan = None

    def _exchange(self, send, send_counts, recv_counts):
        out = torch.empty((int(sum(recv_counts)),) + tuple(send.shape[1:]), dtype=send.dtype, device=send.device)
        if self.world == 1:
            out.copy_(send)
        else:
            dist.all_to_all_single(out, send.contiguous(), output_split_sizes=recv_counts, input_split_sizes=send_counts,
                                   group=self.group)
        return out

    def fetch(self, ids):
        """ids: int32 global row ids (any order, duplicates allowed) -> rows [len(ids), ld] in the same order."""
        ids64 = ids.to(torch.int64)
        owner = torch.searchsorted(self.bounds, ids64, right=True)
        order = torch.argsort(owner, stable=True)
        sorted_ids = ids64[order]
        send_counts = torch.bincount(owner, minlength=self.world)
        recv_counts = torch.empty_like(send_counts)
        if self.world == 1:
            recv_counts.copy_(send_counts)
        else:
            dist.all_to_all_single(recv_counts, send_counts, group=self.group)
        sc, rc = send_counts.tolist(), recv_counts.tolist()
        wanted = self._exchange(sorted_ids, sc, rc)                               # ids other ranks want from me
        local_ids = (wanted - self.lo).to(torch.int32)
        rows_out = self._gather(self.local, local_ids)                            # owners gather
        rows_in = self._exchange(rows_out, rc, sc)                                # rows come back, grouped by owner
        inv = torch.empty_like(order); inv[order] = torch.arange(order.numel(), device=order.device)
        self._plan = (order, sc, rc, local_ids)
        return rows_in[inv]

    def push(self, deltas, target=None):
        """deltas[t] is added to global row ids[t] of the last fetch (duplicates accumulate).  target (optional): a
        tensor partitioned like the table (e.g. this rank's dense GRADIENT shard for an Adam-trained table) that
        receives the rows instead of the table itself."""
        order, sc, rc, local_ids = self._plan
        back = self._exchange(deltas[order].contiguous(), sc, rc)
        dst = self.local if target is None else target
        assert dst.shape[0] == self.hi - self.lo
        self._scatter(dst, local_ids, back)


def sharded_bpr_step(U_local, items: "ShardedTable", tu_local, ti, tj, hyper, bias_col=-1, loss=None, step_fn=None):
    """One BPR step with row-sharded ITEM tables: this rank's triples use local user rows
    (tu_local = local row index) and global item ids; pos/neg item rows are fetched from their owners, the update
    runs locally (ops.bpr_step_rows_f32), the item deltas go back to the owners."""
    from . import ops
    n = tu_local.numel()
    rows = items.fetch(torch.cat([ti, tj]))
    Ri, Rj = rows[:n].contiguous(), rows[n:].contiguous()
    fn = step_fn or ops.bpr_step_rows_f32
    dRi, dRj = fn(U_local, tu_local, Ri, Rj, bias_col, *hyper, loss=loss)
    items.push(torch.cat([dRi, dRj]))


def gather_topk(idx_local, val_local, n_users, group=None):
    """Concatenate user-sharded (idx, val) blocks in rank order (block partition => user order).
    Shards may differ by one row: blocks are padded to the largest shard for the all_gather."""
    world = dist.get_world_size(group)
    sizes = [hi - lo for lo, hi in (shard_range(n_users, r, world) for r in range(world))]
    mx, k = max(sizes), idx_local.shape[1]

    def padded(t, fill):
        if t.shape[0] == mx:
            return t.contiguous()
        out = torch.full((mx, k), fill, dtype=t.dtype, device=t.device)
        out[:t.shape[0]] = t
        return out
    out_i = [torch.empty((mx, k), dtype=idx_local.dtype, device=idx_local.device) for _ in range(world)]
    out_v = [torch.empty((mx, k), dtype=val_local.dtype, device=val_local.device) for _ in range(world)]
    dist.all_gather(out_i, padded(idx_local, -1), group=group)
    dist.all_gather(out_v, padded(val_local, float("-inf")), group=group)
    return (torch.cat([t[:n] for t, n in zip(out_i, sizes)]), torch.cat([t[:n] for t, n in zip(out_v, sizes)]))


# ---------------------------------------------------------------- peer-addressed tables (one NVSwitch box)
def ceil_shard(n_rows, world):
    """Equal shards of ceil(n_rows / world) rows (the last may be short): owner = row // shard_rows — what the
    kernels of csrc/peer.cu compute with one multiply-high."""
    return -(-int(n_rows) // int(world))


class PeerShardedTable:
    """A [n_rows, ld] fp32 table row-sharded over the ranks in memory every rank can address (peer.PeerBuffer).
    `local` = this rank's shard (rows [lo, hi) in its first hi-lo rows), `ptrs` = every shard's device address as
    seen from this process — the `*_shards` argument of the PEER kernels (ops.bpr_step_sampled_peer_f32,
    ops.neumf_gather_peer, ...), which read and atomically update rows in their owner's memory over NVLink: there is
    no fetch/push step and no all-to-all."""

    def __init__(self, n_rows, ld, group=None, device=None, buffer_cls=None):
        from .peer import PeerBuffer
        self.n_rows, self.ld, self.group = int(n_rows), int(ld), group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.shard_rows = ceil_shard(n_rows, self.world)
        self.lo = min(self.rank * self.shard_rows, self.n_rows)
        self.hi = min(self.lo + self.shard_rows, self.n_rows)
        self.buf = (buffer_cls or PeerBuffer)(self.shard_rows * self.ld, group=group, device=device)
        self.local = self.buf.local.view(self.shard_rows, self.ld)
        self.ptrs = self.buf.ptr_array()

    def load_from_full(self, full):
        """Copy this rank's rows out of a full [n_rows, ld] table (init / tests)."""
        self.local[:self.hi - self.lo].copy_(full[self.lo:self.hi])

    def all_rows(self):
        """The whole table gathered on this rank (scoring over sharded item tables): one kernel reading every shard."""
        from . import ops
        ids = torch.arange(self.n_rows, dtype=torch.int32, device=self.local.device)
        return ops.gather_rows_peer_f32(self.ptrs, self.shard_rows, self.ld, ids, self.ld)

    def barrier(self):
        self.buf.barrier()

    def close(self):
        self.local = None
        self.buf.close()


class PeerTableSync:
    """Replicated table (every rank trains on its own copy, which lives in a peer.PeerBuffer) kept consistent by ONE
    kernel per rank and step instead of delta-kernel -> NCCL all-reduce -> apply-kernel: rank r owns the r-th slice,
    reads that slice of every copy over NVLink, agrees on the mean (or sum) of the copies' steps since the slice's
    last agreed value and pushes each copy's correction back with vector atomics (ops.table_reconcile_peer_f32).
    The correction is atomic and per element, so training kernels may run on any copy meanwhile: nothing is lost,
    nothing is counted twice, no rank ever waits for another (ReplicatedTableSync semantics, asynchronously)."""

    def __init__(self, buf, numel=None, reduce="mean", max_ctas=0):
        from . import ops
        assert reduce in ("mean", "sum")
        self.buf, self.ops = buf, ops
        self.numel = int(numel if numel is not None else buf.numel)
        assert self.numel % 4 == 0 and self.numel <= buf.numel
        lo4, hi4 = shard_range(self.numel // 4, buf.rank, buf.world)
        self.lo, self.hi = 4 * lo4, 4 * hi4
        self.scale = 1.0 / buf.world if reduce == "mean" else 1.0
        self.max_ctas = max_ctas
        self.slice_ptrs = buf.ptr_array(self.lo)
        self.prev = buf.local[self.lo:self.hi].clone()

    def reset(self):
        """Collective: re-baseline on the copies' current contents (they must already agree)."""
        self.buf.barrier()
        self.prev.copy_(self.buf.local[self.lo:self.hi])
        self.buf.barrier()

    def sync(self):
        if self.hi > self.lo:
            self.ops.table_reconcile_peer_f32(self.slice_ptrs, self.prev, self.scale, self.max_ctas)

    def flush(self):
        """Collective: after every rank stopped training, one more round makes all copies equal."""
        self.buf.barrier()
        self.sync()
        self.buf.barrier()
