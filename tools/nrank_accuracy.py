"""torchrun --nproc-per-node N tools/nrank_accuracy.py : does the N-rank throughput schedule (users sharded, item table
replicated and reconciled every launch with the AVERAGED deltas — effective item step / N per launch) reach the reference's
nDCG@10?  (VERDICT r1 #2c.)

Same data, split, hyper-parameters and epoch count as tests/golden/bprmf_c1.npz (the reference's own run_experiment on the
ML-1M-shaped file): every rank trains its users' share of the epoch's triples in launches of `--batch` triples with the fused
Hogwild kernel, the replicated item table is reconciled after every launch (peer kernel, or NCCL if peer mapping is
unavailable), after each epoch rank 0 gathers the user shards and computes nDCG@10 with the device metric kernel.
Writes gpurun_out/nrank_accuracy_n{N}.json:  per-epoch nDCG next to the reference's and |final difference|."""
import argparse
import json
import os
import sys
from types import SimpleNamespace

import numpy as np
import pandas as pd
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from elliot_b200 import ops, synth_c1                                                   # noqa: E402
from elliot_b200.dataset import DataSet, eval_csr_of, train_csr_of                      # noqa: E402
from elliot_b200.evaluation import Evaluator                                            # noqa: E402
from elliot_b200.parallel import PeerTableSync, ReplicatedTableSync, shard_range        # noqa: E402
from elliot_b200.run import split_random_subsampling                                    # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=65536, help="triples per launch, summed over the ranks")
ap.add_argument("--seeds", default="42,43,44")
ap.add_argument("--reduce", default="mean", choices=["mean", "sum"],
                help="mean: the ranks' item-table steps are averaged (bench schedule; effective item step / N per launch); "
                     "sum: applied in full (what one GPU would have applied for the same triples)")
args = ap.parse_args()
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local); dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
g = dict(np.load(os.path.join(ROOT, "tests", "golden", "bprmf_c1.npz")))
E, D = int(g["epochs"]), int(g["factors"])
HP = (0.05, 0.0025, 0.0, 0.0025, 0.00025)
u, i, r = synth_c1.rows()
assert synth_c1.checksum(u, i, r) == int(g["checksum"])
df = pd.DataFrame({"userId": u, "itemId": i, "rating": r.astype(float)})
(train, test), = split_random_subsampling(df, 0.2, 42)
cfg = SimpleNamespace(config_test=True, top_k=10, evaluation=SimpleNamespace(simple_metrics=["nDCG"], relevance_threshold=0, cutoffs=[10]))
data = DataSet(cfg, (train, test))
nu, ni, T = data.num_users, data.num_items, data.transactions
indptr, _, srt = train_csr_of(data, dev, set_order=False)
ev = Evaluator(data, SimpleNamespace(meta=SimpleNamespace()))
lo, hi = shard_range(nu, rank, world)
ip_loc = (indptr[lo:hi + 1] - indptr[lo]).contiguous(); ix_loc = srt[int(indptr[lo]):int(indptr[hi])].contiguous()
nl = hi - lo
filt = ops.bloom_build(ip_loc, ix_loc, nl)
share = int(ix_loc.numel())                                   # this rank's triples per epoch (events = transactions, BPRMF.py:119)
out = {"world": world, "reduce": args.reduce, "batch_all_ranks": args.batch, "epochs": E, "reference_ndcg_per_epoch": g["per_epoch"][:, 0].tolist(), "runs": []}
for seed in [int(s) for s in args.seeds.split(",")]:
    rs = np.random.RandomState(seed)
    U0 = rs.normal(0, 0.1, (nu, D)); V0 = rs.normal(0, 0.1, (ni, D))          # BPRMF_model.py:53-56 draw order
    U = torch.from_numpy(U0[lo:hi]).float().to(dev).contiguous()
    n_flat = (ni * D + ni + 3) // 4 * 4
    kind = "single"
    buf = None
    if world > 1:
        try:
            from elliot_b200.peer import PeerBuffer
            buf = PeerBuffer(n_flat, device=dev); kind = "peer:" + buf.kind
        except Exception as e:                                                  # noqa: BLE001
            kind = "nccl (" + str(e)[:60] + ")"
    flat = buf.local if buf is not None else torch.zeros(n_flat, device=dev)
    V = flat[:ni * D].view(ni, D); b = flat[ni * D:ni * D + ni]
    V.copy_(torch.from_numpy(V0).float()); b.zero_()
    sync = None
    if world > 1:
        sync = PeerTableSync(buf, n_flat, reduce=args.reduce) if buf is not None else ReplicatedTableSync([flat], reduce=args.reduce, flat=flat)
        sync.reset()
    per_launch = max(1, args.batch // world)
    drawn, curve = 0, []
    for ep in range(E):
        done = 0
        while done < share:
            n = min(per_launch, share - done)
            ops.bpr_step_sampled_f32(U, V, b, D, nl, ni, ip_loc, ix_loc, n, seed + 1000 * rank, drawn, *HP, filter=filt)
            drawn += n; done += n
            if sync is not None:
                sync.sync()
        if isinstance(sync, PeerTableSync):
            sync.flush()
        # ranks may have different launch counts (shares differ): line up, then evaluate on rank 0
        if world > 1:
            sizes = [shard_range(nu, q, world)[1] - shard_range(nu, q, world)[0] for q in range(world)]
            mx = max(sizes)
            pad = torch.zeros((mx, D), device=dev); pad[:nl] = U
            parts = [torch.empty_like(pad) for _ in range(world)]
            dist.all_gather(parts, pad)
            Ufull = torch.cat([p[:s] for p, s in zip(parts, sizes)])
        else:
            Ufull = U
        if rank == 0:
            idx, _, _ = ops.score_topk_tc(Ufull.contiguous(), V, b, D, 10, indptr, srt, stats=False)
            curve.append(ev.eval_tensors(idx)[10]["test_results"]["nDCG"])
        if world > 1:
            dist.barrier()
    out["runs"].append({"seed": seed, "sync": kind, "ndcg_per_epoch": curve})
    if buf is not None:
        buf.close()
if rank == 0:
    finals = [r_["ndcg_per_epoch"][-1] for r_ in out["runs"]]
    ref = out["reference_ndcg_per_epoch"][-1]
    out.update({"final_mean": float(np.mean(finals)), "reference_final": ref, "abs_diff_of_mean": abs(float(np.mean(finals)) - ref)})
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open(f"gpurun_out/nrank_accuracy_n{world}_{args.reduce}.json", "w"), indent=1)
    print(json.dumps({k: out[k] for k in ("world", "reduce", "final_mean", "reference_final", "abs_diff_of_mean")}))
if world > 1:
    dist.destroy_process_group()
