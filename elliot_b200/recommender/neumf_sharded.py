"""NeuMF with row-sharded embedding tables (BASELINE configs[3]: 10 M users x 1 M items over 4 GPUs).

STATUS (end of round 1): composed only of kernels that are tested on the B200 (`eb_neumf_*`, `eb_gemm_bf16_tn`,
`eb_gather/scatter_add_rows_f32`, `eb_adam_dense_f32`) and of `parallel.ShardedTable` / `GradAllReduce` (gloo-tested).
Run on one B200 at world = 1 it tracks the ordinary model to 1e-8 over 3 steps (`tools/neumf_sharded_w1.py`,
profiles/r1f_neumf_sharded_w1.json); the multi-GPU run (`tools/neumf_sharded_check.py`) is pending — the round's GPU
budget was spent.  Nothing imports this module by default.

Layout: users are block-partitioned over the ranks (`shard_range`), each rank holds the MF and MLP rows of its own
users and samples only for them, so user rows never move.  The two item tables are stored side by side in ONE
[items, 2f] table (MF | MLP) that is block-partitioned over the ranks: one all-to-all of ids and one of rows fetches
both rows of every sampled item (`ShardedTable.fetch`), the step runs on the fetched copies with the ordinary NeuMF
kernels (tables = the B fetched rows, item index = position), and the per-sample item-row gradients go back to the
owners' dense gradient shard (`push(target=)`), where Keras Adam runs over the shard — the same dense-over-all-rows
Adam as the reference (neural_matrix_factorization_model.py:72,98-104) because every row lives on exactly one rank.
The MLP / head weights are replicated and their gradients averaged (`GradAllReduce`).  BinaryCrossentropy is a batch
mean: each rank normalises by its local batch, so embedding gradients are `world` x the global-batch gradient — Adam's
update is invariant to that scale up to epsilon.
"""
import torch
import torch.distributed as dist

from .. import ops
from ..parallel import GradAllReduce, ShardedTable, shard_range
from .neumf import NeuralMatrixFactorizationModel

_MLP_KEYS = ("W1", "b1", "W2", "b2", "W3", "b3", "wp", "bp")


class ShardedNeuMFModel(NeuralMatrixFactorizationModel):
    def __init__(self, num_users, num_items, f, learning_rate, random_seed, device, group=None):
        # same initial weights as the single-GPU model with the same seed (every rank draws the full tables and keeps
        # its blocks: init only, at C4 scale 2.6 GB of scratch per table)
        super().__init__(num_users, num_items, f, learning_rate, random_seed, device)
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.ulo, self.uhi = shard_range(num_users, self.rank, self.world)
        self.ilo, self.ihi = shard_range(num_items, self.rank, self.world)
        P = self.P
        items = torch.cat([P["I_mf"][self.ilo:self.ihi], P["I_mlp"][self.ilo:self.ihi]], dim=1).contiguous()   # [n_il, 2f]
        P["U_mf"] = P["U_mf"][self.ulo:self.uhi].contiguous(); P["U_mlp"] = P["U_mlp"][self.ulo:self.uhi].contiguous()
        del P["I_mf"], P["I_mlp"]
        P["I"] = items
        zl = lambda t: torch.zeros_like(t)
        # MLP / head gradients in one flat buffer (one all-reduce); embedding gradients per table
        mlp_n = sum(P[k].numel() for k in _MLP_KEYS)
        self._mlp_flat = torch.zeros(mlp_n, device=self.device)
        self.G, off = {}, 0
        for k in _MLP_KEYS:
            self.G[k] = self._mlp_flat[off:off + P[k].numel()].view_as(P[k]); off += P[k].numel()
        for k in ("U_mf", "U_mlp", "I"):
            self.G[k] = zl(P[k])
        self.M = {k: zl(v) for k, v in P.items()}
        self.V = {k: zl(v) for k, v in P.items()}
        self.items = ShardedTable(num_items, P["I"], group)
        self.dp = GradAllReduce(self._mlp_flat, group, extra=self._loss)

    def train_step(self, batch):
        """batch = (LOCAL user row int32, GLOBAL item id int32, label float32) for users this rank owns."""
        u, it, y = batch
        f, B, P, G = self.f, u.numel(), self.P, self.G
        dev = self.device
        rows = self.items.fetch(it)                                              # [B, 2f] copies of (MF | MLP) item rows
        R_mf, R_mlp = rows[:, :f].contiguous(), rows[:, f:].contiguous()
        pos = torch.arange(B, dtype=torch.int32, device=dev)                      # item "index" = position among the copies
        x0 = torch.empty((B, 2 * f), device=dev); pm = torch.empty((B, f), device=dev)
        ops.neumf_gather(P["U_mf"], R_mf, P["U_mlp"], R_mlp, f, u, pos, x0, pm)
        h1, h2, h3 = self._mlp(x0)
        dpm = torch.empty_like(pm); dpre3 = torch.empty_like(h3)
        self._loss.zero_()
        ops.neumf_head(pm, h3, f, P["wp"], P["bp"], label=y, dpm=dpm, dh3=dpre3, dwp=G["wp"], dbp=G["bp"], loss=self._loss)
        T = lambda t: ops.to_bf16(t, transpose=True)
        ops.gemm_bf16_tn(T(dpre3), T(h2), f, 2 * f, B, out=G["W3"]); ops.colsum(dpre3, G["b3"])
        dpre2 = ops.relu_bwd(ops.gemm_bf16_tn(ops.to_bf16(dpre3), self.Wt["W3"], B, 2 * f, f), h2)
        ops.gemm_bf16_tn(T(dpre2), T(h1), 2 * f, 4 * f, B, out=G["W2"]); ops.colsum(dpre2, G["b2"])
        dpre1 = ops.relu_bwd(ops.gemm_bf16_tn(ops.to_bf16(dpre2), self.Wt["W2"], B, 4 * f, 2 * f), h1)
        ops.gemm_bf16_tn(T(dpre1), T(x0), 4 * f, 2 * f, B, out=G["W1"]); ops.colsum(dpre1, G["b1"])
        dx0 = ops.gemm_bf16_tn(ops.to_bf16(dpre1), self.Wt["W1"], B, 2 * f, 4 * f)
        dR_mf = torch.zeros((B, f), device=dev); dR_mlp = torch.zeros((B, f), device=dev)
        ops.neumf_scatter(P["U_mf"], R_mf, f, u, pos, dpm, dx0, G["U_mf"], dR_mf, G["U_mlp"], dR_mlp)
        self.items.push(torch.cat([dR_mf, dR_mlp], dim=1), target=G["I"])         # item-row grads -> owners' dense grad shard
        self.dp.sync()                                                            # MLP grads averaged, loss summed
        self.step += 1
        for k in P:
            ops.adam_dense_f32(P[k], self.M[k], self.V[k], G[k], self.lr, self.step)
        self._refresh()
        return self._loss

    def get_recs_topk(self, *a, **k):
        raise NotImplementedError("scoring over sharded NeuMF tables: gather the item shard per user block (next round)")
