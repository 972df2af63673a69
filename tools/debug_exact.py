import numpy as np, torch, oracle
from elliot_b200 import ops
g = dict(np.load("tests/golden/bprmf_tiny.npz")); d = int(g["d"]); hp = [float(x) for x in g["hp"]]
DEV = "cuda:0"
for k in [1, 2, 3, 5, 10, 20, 50, 100, 200, 420]:
    U, V, b = g["U0"].copy(), g["V0"].copy(), np.zeros(len(g["items"]))
    oracle.bpr_update_seq(U, V, b, g["tu"][:k], g["ti"][:k], g["tj"][:k], *hp)
    Ud = torch.from_numpy(g["U0"].copy()).to(DEV); Vd = torch.from_numpy(g["V0"].copy()).to(DEV)
    bd = torch.zeros(len(g["items"]), dtype=torch.float64, device=DEV)
    ops.bpr_exact_f64(Ud, Vd, bd, d, torch.from_numpy(g["tu"][:k].copy()).to(DEV), torch.from_numpy(g["ti"][:k].copy()).to(DEV),
                      torch.from_numpy(g["tj"][:k].copy()).to(DEV), *hp)
    torch.cuda.synchronize()
    print(k, np.abs(Ud.cpu().numpy() - U).max(), np.abs(Vd.cpu().numpy() - V).max(), np.abs(bd.cpu().numpy() - b).max(),
          "triples", list(zip(g["tu"][:min(k,3)], g["ti"][:min(k,3)], g["tj"][:min(k,3)])))
