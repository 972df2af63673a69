"""Single-GPU smoke of ShardedNeuMFModel (world = 1: every all-to-all degenerates to a copy): must track the ordinary
model step for step.  Writes gpurun_out/neumf_sharded_w1.json."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from elliot_b200.recommender.neumf import NeuralMatrixFactorizationModel
from elliot_b200.recommender.neumf_sharded import ShardedNeuMFModel
dev = torch.device("cuda:0")
NU, NI, F, B, STEPS = 5000, 3000, 64, 4096, 3
sh = ShardedNeuMFModel(NU, NI, F, 1e-3, 42, dev); ref = NeuralMatrixFactorizationModel(NU, NI, F, 1e-3, 42, dev)
g = torch.Generator(device=dev); g.manual_seed(1)
losses = []
for s in range(STEPS):
    u = torch.randint(0, NU, (B,), device=dev, generator=g, dtype=torch.int32)
    it = torch.randint(0, NI, (B,), device=dev, generator=g, dtype=torch.int32)
    y = (torch.rand(B, device=dev, generator=g) < 0.3).float()
    losses.append((float(sh.train_step((u, it, y)).item()), float(ref.train_step((u, it, y)).item())))
rel = lambda a, b: float((a - b).norm() / b.norm().clamp_min(1e-12))
d = {"U_mf": rel(sh.P["U_mf"], ref.P["U_mf"]), "U_mlp": rel(sh.P["U_mlp"], ref.P["U_mlp"]),
     "I_mf": rel(sh.P["I"][:, :F], ref.P["I_mf"]), "I_mlp": rel(sh.P["I"][:, F:], ref.P["I_mlp"]),
     **{k: rel(sh.P[k], ref.P[k]) for k in ("W1", "W2", "W3", "wp")}}
out = {"losses": losses, "rel_diff": d, "ok": max(d.values()) < 1e-5}
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/neumf_sharded_w1.json", "w"), indent=1)
print(json.dumps(out))
