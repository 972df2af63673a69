"""2+ GPU check of data-parallel MultiVAE: torchrun --nproc-per-node N tools/vae_dp_check.py
Every rank builds the same model, trains STEPS steps on its slice of each batch with gradients averaged by
GradAllReduce, and rank 0 compares the weights with a single-process emulation of the same schedule (the slices'
gradients computed one after the other with the ranks' noise salts, averaged, one Adam step).  Writes
gpurun_out/vae_dp_check.json."""
import json
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from elliot_b200 import ops                                                     # noqa: E402
from elliot_b200.parallel import shard_range                                    # noqa: E402
from elliot_b200.recommender.multi_vae import VariationalAutoEncoder            # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
U, I, H, L, B, STEPS = 20000, 26744, 600, 200, 512 * world, 6
g = torch.Generator(device=dev); g.manual_seed(1)
cand, _ = torch.sort((torch.rand(U, 60, device=dev, generator=g) ** 2 * I).to(torch.int32).clamp_(max=I - 1), dim=1)
keep = torch.ones_like(cand, dtype=torch.bool); keep[:, 1:] = cand[:, 1:] != cand[:, :-1]
indptr = torch.zeros(U + 1, dtype=torch.int64, device=dev); indptr[1:] = torch.cumsum(keep.sum(1), 0)
indices = cand[keep].contiguous()


def make():
    return VariationalAutoEncoder(I, H, L, 1e-3, 0.5, 0.01, 42, indptr, indices, dev)


order = torch.randperm(U, device=dev, generator=g).to(torch.int32)
dp = make(); dp.enable_data_parallel()
torch.cuda.synchronize(); dist.barrier()
losses, t0 = [], time.perf_counter()
for s in range(STEPS):
    rows = order[s * B:(s + 1) * B]
    lo, hi = shard_range(B, rank, world)
    losses.append(dp.train_step(rows[lo:hi].contiguous(), 0.1))
torch.cuda.synchronize()
dp_ms = (time.perf_counter() - t0) * 1e3 / STEPS
same = torch.tensor([float(dp.P["W4"].double().sum().item())], device=dev, dtype=torch.float64)
lst = [torch.zeros_like(same) for _ in range(world)]
dist.all_gather(lst, same)
replicas_agree = all(abs(x.item() - lst[0].item()) < 1e-9 * max(1.0, abs(lst[0].item())) for x in lst)

out = {"world": world, "replicas_agree": replicas_agree, "dp_ms_per_step": dp_ms, "global_batch": B, "loss": losses}
if rank == 0:
    # emulation: same schedule on one GPU — per slice: gradients with that rank's salt, accumulated, averaged, one Adam step
    em = make()
    for s in range(STEPS):
        rows = order[s * B:(s + 1) * B]
        acc = torch.zeros_like(em._gflat)
        em.step += 1
        for r in range(world):
            lo, hi = shard_range(B, r, world)
            em._salt = 0x9E3779B1 * r
            em._acc.zero_()
            em.compute_grads(rows[lo:hi].contiguous(), 0.1, em.step)
            acc += em._gflat; em._gflat.zero_()
        em._gflat.copy_(acc / world)
        em.apply_grads()
    diffs = {k: float((em.P[k] - dp.P[k]).abs().max().item()) for k in em.P}
    scale = {k: float(dp.P[k].abs().max().item()) for k in em.P}
    out.update({"max_abs_diff_vs_emulation": diffs, "max_abs_weight": scale,
                "ok": bool(replicas_agree and all(diffs[k] <= 2e-4 * max(scale[k], 1e-3) + 2e-6 for k in diffs))})
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/vae_dp_check.json", "w"), indent=1)
    print(json.dumps(out))
dist.barrier()
dist.destroy_process_group()
