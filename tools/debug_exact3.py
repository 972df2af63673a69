import numpy as np, torch, oracle
from elliot_b200 import ops
g = dict(np.load("tests/golden/bprmf_tiny.npz")); d = int(g["d"]); hp = [float(x) for x in g["hp"]]
DEV = "cuda:0"
for k in list(range(10, 21)):
  for rep in range(3):
    U, V, b = g["U0"].copy(), g["V0"].copy(), np.zeros(len(g["items"]))
    oracle.bpr_update_seq(U, V, b, g["tu"][:k], g["ti"][:k], g["tj"][:k], *hp)
    Ud = torch.from_numpy(g["U0"].copy()).to(DEV); Vd = torch.from_numpy(g["V0"].copy()).to(DEV)
    bd = torch.zeros(len(g["items"]), dtype=torch.float64, device=DEV)
    ops.bpr_exact_f64(Ud, Vd, bd, d, torch.from_numpy(g["tu"][:k].copy()).to(DEV), torch.from_numpy(g["ti"][:k].copy()).to(DEV),
                      torch.from_numpy(g["tj"][:k].copy()).to(DEV), *hp)
    torch.cuda.synchronize()
    du = np.abs(Ud.cpu().numpy() - U).max(1); dv = np.abs(Vd.cpu().numpy() - V).max(1); db = np.abs(bd.cpu().numpy() - b)
    print(k, rep, "U rows", {int(r): float(f"{du[r]:.2e}") for r in np.nonzero(du > 1e-13)[0]}, "V rows", {int(r): float(f"{dv[r]:.2e}") for r in np.nonzero(dv > 1e-13)[0]},
          "b", {int(r): float(f"{db[r]:.2e}") for r in np.nonzero(db > 1e-13)[0]}, "last", (g["tu"][k-1], g["ti"][k-1], g["tj"][k-1]))
