import numpy as np, torch, oracle
from elliot_b200 import ops
np.set_printoptions(precision=6, linewidth=200)
g = dict(np.load("tests/golden/bprmf_tiny.npz")); d = int(g["d"]); hp = [float(x) for x in g["hp"]]
DEV = "cuda:0"
k = 10
for rep in range(4):
    U, V, b = g["U0"].copy(), g["V0"].copy(), np.zeros(len(g["items"]))
    oracle.bpr_update_seq(U, V, b, g["tu"][:k], g["ti"][:k], g["tj"][:k], *hp)
    Ud = torch.from_numpy(g["U0"].copy()).to(DEV); Vd = torch.from_numpy(g["V0"].copy()).to(DEV)
    bd = torch.zeros(len(g["items"]), dtype=torch.float64, device=DEV)
    ops.bpr_exact_f64(Ud, Vd, bd, d, torch.from_numpy(g["tu"][:k].copy()).to(DEV), torch.from_numpy(g["ti"][:k].copy()).to(DEV),
                      torch.from_numpy(g["tj"][:k].copy()).to(DEV), *hp)
    torch.cuda.synchronize()
    print("rep", rep)
    print(" dU23", Ud.cpu().numpy()[23] - U[23])
    print(" dV39", Vd.cpu().numpy()[39] - V[39])
    print(" dV29", Vd.cpu().numpy()[29] - V[29])
    print(" db39", bd.cpu().numpy()[39] - b[39], "db29", bd.cpu().numpy()[29] - b[29])
