"""2+ GPU check of NeuMF over row-sharded tables (BASELINE configs[3]) — run FIRST next round:
    python -m torch.distributed.run --nproc-per-node 2 tools/neumf_sharded_check.py
Each rank trains STEPS steps on samples of the users it owns; rank 0 then replays the same global batches (the
concatenation of all ranks' samples) on an ordinary single-GPU NeuralMatrixFactorizationModel with the same seed
and compares the embedding rows it owns and the MLP weights.  The per-rank BCE means make the sharded embedding
gradients `world` x the global-batch ones (Adam-invariant up to epsilon), so tolerances are loose (1e-3 relative)."""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from elliot_b200.parallel import shard_range                                            # noqa: E402
from elliot_b200.recommender.neumf import NeuralMatrixFactorizationModel               # noqa: E402
from elliot_b200.recommender.neumf_sharded import ShardedNeuMFModel                     # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
NU, NI, F, B, STEPS = 40000, 20000, 64, 8192, 4
sh = ShardedNeuMFModel(NU, NI, F, 1e-3, 42, dev)
g = torch.Generator(device=dev); g.manual_seed(100 + rank)
ulo, uhi = shard_range(NU, rank, world)
batches = []
for s in range(STEPS):
    u = torch.randint(0, uhi - ulo, (B,), device=dev, generator=g, dtype=torch.int32)
    it = torch.randint(0, NI, (B,), device=dev, generator=g, dtype=torch.int32)
    y = (torch.rand(B, device=dev, generator=g) < 0.3).float()
    batches.append((u, it, y))
    sh.train_step((u, it, y))
torch.cuda.synchronize()
# gather every rank's samples (global user ids) on all ranks
allb = []
for (u, it, y) in batches:
    parts = [[torch.empty_like(t) for _ in range(world)] for t in (u, it, y)]
    for p, t in zip(parts, ((u + ulo).contiguous(), it, y)):
        dist.all_gather(p, t)
    allb.append(tuple(torch.cat(p) for p in parts))
out = {"world": world, "finite": bool(torch.isfinite(sh.P["I"]).all().item())}
if rank == 0:
    ref = NeuralMatrixFactorizationModel(NU, NI, F, 1e-3, 42, dev)
    for b in allb:
        ref.train_step(b)
    f = F
    ilo, ihi = shard_range(NI, rank, world)
    rel = lambda a, b: float((a - b).norm() / b.norm().clamp_min(1e-12))
    diffs = {"U_mf": rel(sh.P["U_mf"], ref.P["U_mf"][ulo:uhi]), "U_mlp": rel(sh.P["U_mlp"], ref.P["U_mlp"][ulo:uhi]),
             "I_mf": rel(sh.P["I"][:, :f], ref.P["I_mf"][ilo:ihi]), "I_mlp": rel(sh.P["I"][:, f:], ref.P["I_mlp"][ilo:ihi]),
             **{k: rel(sh.P[k], ref.P[k]) for k in ("W1", "W2", "W3", "wp")}}
    out.update({"rel_diff_vs_single_gpu": diffs, "ok": bool(out["finite"] and max(diffs.values()) < 1e-3)})
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/neumf_sharded_check.json", "w"), indent=1)
    print(json.dumps(out))
dist.barrier()
dist.destroy_process_group()
