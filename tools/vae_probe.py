"""MultiVAE step time at the C3 shape (ML-20M: 138493 x 26744, batch 512, 600/200)."""
import json, os, sys
import numpy as np, torch, time
from elliot_b200.recommender.multi_vae import VariationalAutoEncoder
dev = "cuda:0"
nu, ni, B = 138493, 26744, 512
g = torch.Generator(device=dev); g.manual_seed(0)
per = 144
cand = (torch.rand(nu, per, device=dev, generator=g) ** 2 * ni).to(torch.int32).clamp_(max=ni - 1)
cand, _ = torch.sort(cand, dim=1); keep = torch.ones_like(cand, dtype=torch.bool); keep[:, 1:] = cand[:, 1:] != cand[:, :-1]
indptr = torch.zeros(nu + 1, dtype=torch.int64, device=dev); indptr[1:] = torch.cumsum(keep.sum(1), 0); indices = cand[keep].contiguous()
out = {}
FAST = os.environ.get("VAE_PROBE_FAST") == "1"      # under ncu: native path only, few steps
for native in ((True,) if FAST else (True, False)):
    m = VariationalAutoEncoder(ni, 600, 200, 1e-3, 0.0, 0.01, 42, indptr, indices, dev)
    m.native = native
    rows = torch.randperm(nu, device=dev, generator=g)[:B].to(torch.int32)
    for _ in range(1 if FAST else 3): m.train_step(rows, 0.1)
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(2 if FAST else 20): loss = m.train_step(rows, 0.1)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / (2 if FAST else 20)
    # the step without the per-step loss read-back (.tolist() synchronises): phases issued back to back
    m._acc.zero_()
    e0.record()
    for _ in range(0 if FAST else 20):
        m.step += 1
        m.compute_grads(rows, 0.1, m.step); m.apply_grads()
    e1.record(); torch.cuda.synchronize()
    ms_async = e0.elapsed_time(e1) / 20
    flop = 3 * 2 * B * ni * 600 + 3 * 2 * B * 600 * 400 + 3 * 2 * B * 200 * 600     # fwd+bwd dense (input layer is a gather)
    key = "native" if native else "python_sequence"
    out[key] = {"ms_per_step": ms, "users_per_s": B / ms * 1e3, "ms_per_step_no_readback": ms_async,
                "users_per_s_no_readback": B / ms_async * 1e3, "dense_equiv_tflops_no_readback": flop / ms_async / 1e9, "loss": loss}
    print(key, out[key])
os.makedirs("gpurun_out", exist_ok=True)
json.dump({"workload": "MultiVAE C3 shape: 138493 x 26744, batch 512, 600/200, dropout 0", **out}, open("gpurun_out/vae_probe.json", "w"), indent=1)
