"""NeuMF with row-sharded embedding tables (BASELINE configs[3]: 10 M users x 1 M items over 4 GPUs).

Layout.  Users are block-partitioned over the ranks (`shard_range`): a rank holds the MF and MLP rows of its own
users and samples only for them, so user rows never move.  The two ITEM tables live side by side in ONE
[items, 2f] table (row = [I_mf | I_mlp]) that is row-sharded over the ranks in peer-addressable memory
(`parallel.PeerShardedTable`), and so does its dense gradient.  The step (neural_matrix_factorization_model.py:74-106):

    eb_neumf_gather_peer    item rows are READ from their owner's memory over NVLink inside the gather kernel
    MLP forward / head / MLP backward      (tensor cores, local — as the single-GPU model)
    eb_neumf_scatter_peer   item-row gradients are ADDED into the owner's gradient shard over NVLink (vector atomics)
    all-reduce of the MLP / head gradients + loss (one flat buffer, NCCL)  — also the point after which every
                            rank's scatter has landed
    dense Keras Adam on everything this rank owns: its user rows, its item shard, its MLP replica
    one-element all-reduce  — nobody gathers item rows of the next step before every owner's Adam is done

There is no all-to-all and no index bookkeeping: the exchange is the loads and atomics of the two kernels themselves.
Keras Adam over all rows of every table each step is the reference's semantics and, because every
row has exactly one owner, it stays exact under sharding.  BinaryCrossentropy is a batch mean: with `global_mean`
(default) each rank normalises by the GLOBAL batch size, so embedding gradients equal the single-GPU ones; the MLP
gradients are summed over the ranks.  Checked against the single-GPU model by tools/neumf_sharded_check.py (2 and 4
GPUs) and, with several shards inside one GPU, by tests/test_gpu_peer.py.
"""
import math

import torch
import torch.distributed as dist

from .. import ops
from ..parallel import PeerShardedTable, shard_range

_MLP_KEYS = ("W1", "b1", "W2", "b2", "W3", "b3", "wp", "bp")


class ShardedNeuMFModel:
    def __init__(self, num_users, num_items, f, learning_rate, random_seed, device, group=None, full_init=True):
        """full_init=True: every rank draws the full tables with the single-GPU model's stream and keeps its blocks
        (bit-identical start, for checks; needs room for the full tables once).  full_init=False: each rank draws only
        its own rows from a (seed, rank) stream (C4 scale)."""
        assert f % 8 == 0 and 8 <= f <= 128
        self.nu, self.ni, self.f, self.lr = num_users, num_items, f, learning_rate
        self.device = torch.device(device)
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.ulo, self.uhi = shard_range(num_users, self.rank, self.world)
        dev = self.device
        self.items = PeerShardedTable(num_items, 2 * f, group, dev)         # [shard_rows, 2f]: I_mf | I_mlp
        self.items_grad = PeerShardedTable(num_items, 2 * f, group, dev)    # owners' dense gradient shards
        ilo, ihi = self.items.lo, self.items.hi
        g = torch.Generator(device=dev)

        def glorot(rows, cols, fan_in, fan_out):                              # GlorotUniform (:38)
            lim = math.sqrt(6.0 / (fan_in + fan_out))
            return (torch.rand((rows, cols), device=dev, generator=g) * 2 - 1) * lim
        z = lambda n: torch.zeros((n + 3) // 4 * 4, device=dev)
        P = {}
        if full_init:                                                         # same draw order as NeuralMatrixFactorizationModel
            g.manual_seed(int(random_seed))
            t = glorot(num_users, f, num_users, f); P["U_mf"] = t[self.ulo:self.uhi].clone(); del t
            t = glorot(num_items, f, num_items, f); self.items.local[:ihi - ilo, :f].copy_(t[ilo:ihi]); del t
            t = glorot(num_users, f, num_users, f); P["U_mlp"] = t[self.ulo:self.uhi].clone(); del t
            t = glorot(num_items, f, num_items, f); self.items.local[:ihi - ilo, f:].copy_(t[ilo:ihi]); del t
        else:
            g.manual_seed(int(random_seed) * 1000003 + 17 * self.rank + 1)
            nl = self.uhi - self.ulo
            P["U_mf"] = glorot(nl, f, num_users, f); P["U_mlp"] = glorot(nl, f, num_users, f)
            self.items.local[:ihi - ilo].copy_(glorot(ihi - ilo, 2 * f, num_items, f))
            g.manual_seed(int(random_seed))                                   # MLP replicas must start equal
        P.update({"W1": glorot(4 * f, 2 * f, 2 * f, 4 * f), "b1": z(4 * f),
                  "W2": glorot(2 * f, 4 * f, 4 * f, 2 * f), "b2": z(2 * f),
                  "W3": glorot(f, 2 * f, 2 * f, f), "b3": z(f),
                  "wp": glorot(1, 2 * f, 2 * f, 1).reshape(-1).contiguous(), "bp": z(1)})
        P["I"] = self.items.local                                             # this rank's item shard (peer-visible)
        self.P = P
        zl = lambda t: torch.zeros_like(t)
        # MLP / head gradients + the loss accumulator share one flat buffer -> ONE all-reduce per step
        mlp_n = sum(P[k].numel() for k in _MLP_KEYS)
        self._flat = torch.zeros(mlp_n + 4, device=dev)
        self.G, off = {}, 0
        for k in _MLP_KEYS:
            self.G[k] = self._flat[off:off + P[k].numel()].view_as(P[k]); off += P[k].numel()
        self._loss_f32 = self._flat[mlp_n:mlp_n + 1]
        self._mlp_n = mlp_n
        self.G["U_mf"] = zl(P["U_mf"]); self.G["U_mlp"] = zl(P["U_mlp"]); self.G["I"] = self.items_grad.local
        self.M = {k: zl(v) for k, v in P.items()}
        self.V = {k: zl(v) for k, v in P.items()}
        self.step = 0
        self._loss = torch.zeros(1, dtype=torch.float64, device=dev)
        self._tick = torch.zeros(1, device=dev)
        self.global_mean = True
        self._refresh()
        self.items.barrier()                                                  # every shard initialised before anyone reads

    def _refresh(self):
        P = self.P
        self.Wb = {k: ops.to_bf16(P[k]) for k in ("W1", "W2", "W3")}

    def _mlp(self, x0):
        f, B = self.f, x0.shape[0]
        x0b = ops.to_bf16(x0)
        # each layer's epilogue writes the bf16 operand copy of its output beside the fp32 activations
        h1, h1b = ops.gemm_bf16_tn(x0b, self.Wb["W1"], B, 4 * f, 2 * f, bias=self.P["b1"], act=2, out_bf16=True)
        h2, h2b = ops.gemm_bf16_tn(h1b, self.Wb["W2"], B, 2 * f, 4 * f, bias=self.P["b2"], act=2, out_bf16=True)
        h3 = ops.gemm_bf16_tn(h2b, self.Wb["W3"], B, f, 2 * f, bias=self.P["b3"], act=2)
        self._act_b = (x0b, h1b, h2b)                     # row-major bf16 copies, read again ("rows are K") by the weight-gradient GEMMs
        return h1, h2, h3

    def train_step(self, batch, global_batch=None):
        """batch = (LOCAL user row int32, GLOBAL item id int32, label float32) for users this rank owns.
        global_batch: number of samples of all ranks in this step (default: world * local batch)."""
        u, it, y = batch
        f, B, P, G = self.f, u.numel(), self.P, self.G
        dev, I, GI = self.device, self.items, self.items_grad
        if self.world > 4:
            # rows from more than four peers: the gather and the scatter see the batch owner by owner (12x the NVLink row rate
            # on 8 GPUs, ops.group_by_owner); the samples of a step are exchangeable, the loss and the gradients are sums
            u, it, y = ops.group_by_owner([u, it, y], 1, -1, I.shard_rows, self.rank, self.world)
        x0 = torch.empty((B, 2 * f), device=dev); pm = torch.empty((B, f), device=dev)
        ops.neumf_gather_peer(P["U_mf"], P["U_mlp"], I.ptrs, I.shard_rows, 2 * f, f, u, it, x0, pm)
        h1, h2, h3 = self._mlp(x0)
        dpm = torch.empty_like(pm); dpre3 = torch.empty_like(h3)
        self._loss.zero_()
        gb = (global_batch if global_batch is not None else B * self.world) if self.global_mean else B
        ops.neumf_head(pm, h3, f, P["wp"], P["bp"], label=y, dpm=dpm, dh3=dpre3, dwp=G["wp"], dbp=G["bp"], loss=self._loss,
                       mean_over=gb)
        # backward: dW = dY^T . X contracts over the batch rows of both row-major operands, dX = dY . W reads the [out][in] kernel as
        # a [K][N] matrix — "rows are K" operands of eb_gemm_bf16, no transposed copies of activations or weights
        x0b, h1b, h2b = self._act_b
        d3b = ops.to_bf16(dpre3)
        ops.gemm_bf16(d3b, h2b, f, 2 * f, B, a_rows_are_k=True, b_rows_are_k=True, out=G["W3"]); ops.colsum(dpre3, G["b3"])
        dpre2, d2b = ops.relu_bwd(ops.gemm_bf16(d3b, self.Wb["W3"], B, 2 * f, f, b_rows_are_k=True), h2, copy_bf16=True)
        ops.gemm_bf16(d2b, h1b, 2 * f, 4 * f, B, a_rows_are_k=True, b_rows_are_k=True, out=G["W2"]); ops.colsum(dpre2, G["b2"])
        dpre1, d1b = ops.relu_bwd(ops.gemm_bf16(d2b, self.Wb["W2"], B, 4 * f, 2 * f, b_rows_are_k=True), h1, copy_bf16=True)
        ops.gemm_bf16(d1b, x0b, 4 * f, 2 * f, B, a_rows_are_k=True, b_rows_are_k=True, out=G["W1"]); ops.colsum(dpre1, G["b1"])
        dx0 = ops.gemm_bf16(d1b, self.Wb["W1"], B, 2 * f, 4 * f, b_rows_are_k=True)
        ops.neumf_scatter_peer(P["U_mf"], I.ptrs, GI.ptrs, I.shard_rows, 2 * f, f, u, it, dpm, dx0, G["U_mf"], G["U_mlp"])
        if self.world > 1:
            self._loss_f32.copy_(self._loss.to(torch.float32))
            dist.all_reduce(self._flat, group=self.group)            # MLP grads + loss summed; every scatter has landed after it
            if not self.global_mean:
                self._flat[:self._mlp_n].div_(self.world)
            self._loss.copy_(self._loss_f32.to(torch.float64))
        self.step += 1
        for k in P:
            ops.adam_dense_f32(P[k], self.M[k], self.V[k], G[k], self.lr, self.step)
        if self.world > 1:
            dist.all_reduce(self._tick, group=self.group)            # owners' Adam done before the next step's gathers
        self._refresh()
        return self._loss

    def full_item_tables(self):
        """(I_mf, I_mlp) of ALL items on this rank — for scoring: one kernel reads every owner's shard over NVLink."""
        rows = self.items.all_rows()
        f = self.f
        return rows[:, :f].contiguous(), rows[:, f:].contiguous()

    def get_recs_topk(self, u0, u1, k, mask_indptr, mask_indices):
        """sigmoid outputs for LOCAL users [u0, u1) x all items -> masked top-k (get_recs/get_top_k, :119-148): users are
        sharded (no collective in the scoring itself), the item tables are gathered once per model step."""
        f, P, ni, nb = self.f, self.P, self.ni, u1 - u0
        if getattr(self, "_items_step", None) != self.step:
            self._Imf, Imlp = self.full_item_tables()
            self._Ai = ops.gemm_bf16_tn(ops.to_bf16(Imlp), self.Wb["W1"][:, f:2 * f], ni, 4 * f, f)
            self._items_step = self.step
        Au = ops.gemm_bf16_tn(ops.to_bf16(P["U_mlp"][u0:u1]), self.Wb["W1"][:, :f], nb, 4 * f, f)
        pairs = nb * ni
        h1 = torch.empty((pairs, 4 * f), dtype=self.Wb["W2"].dtype, device=self.device)       # bf16 (fp32 in checking mode)
        ops.neumf_pair_h1(Au, self._Ai, P["b1"], nb, ni, 4 * f, h1)
        h2 = ops.gemm_bf16_tn(h1, self.Wb["W2"], pairs, 2 * f, 4 * f, bias=P["b2"], act=2)
        h3 = ops.gemm_bf16_tn(ops.to_bf16(h2), self.Wb["W3"], pairs, f, 2 * f, bias=P["b3"], act=2)
        prob = torch.empty((nb, ni), device=self.device)
        ops.neumf_pair_head(P["U_mf"], self._Imf, f, u0, nb, ni, h3, P["wp"], P["bp"], prob)
        rows = torch.arange(u0, u1, dtype=torch.int32, device=self.device)
        return ops.dense_topk(prob, k, mask_indptr, mask_indices, rows)

    def close(self):
        self.P.pop("I", None); self.G.pop("I", None)
        self.items.close(); self.items_grad.close()
