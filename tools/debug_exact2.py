import numpy as np, torch, oracle
from elliot_b200 import ops
g = dict(np.load("tests/golden/bprmf_tiny.npz")); d = int(g["d"]); hp = [float(x) for x in g["hp"]]
DEV = "cuda:0"
k = 50
Ud = torch.from_numpy(g["U0"].copy()).to(DEV); Vd = torch.from_numpy(g["V0"].copy()).to(DEV)
bd = torch.zeros(len(g["items"]), dtype=torch.float64, device=DEV)
tu, ti, tj = g["tu"][:k].copy(), g["ti"][:k].copy(), g["tj"][:k].copy()
ops.bpr_exact_f64(Ud, Vd, bd, d, torch.from_numpy(tu).to(DEV), torch.from_numpy(ti).to(DEV), torch.from_numpy(tj).to(DEV), *hp)
torch.cuda.synchronize()
ws = ops._ws_exact.buf.cpu().numpy()
al = lambda x: (x + 255) // 256 * 256
o_in = 0; o_out = al(8 * 3 * k); o_ku = o_out + al(8 * 3 * k); o_ki = o_ku + al(4 * k); o_kj = o_ki + al(4 * k); o_cnt = o_kj + al(4 * k)
keys_in = ws[o_in:o_in + 8 * 3 * k].view(np.uint64); keys_out = ws[o_out:o_out + 8 * 3 * k].view(np.uint64)
ku = ws[o_ku:o_ku + 4 * k].view(np.int32); ki = ws[o_ki:o_ki + 4 * k].view(np.int32); kj = ws[o_kj:o_kj + 4 * k].view(np.int32)
print("keys_in U ", keys_in[:8]); print("keys_out U", keys_out[:8]); print("sorted U ok", np.array_equal(np.sort(keys_in[:k]), keys_out[:k]))
print("sorted I ok", np.array_equal(np.sort(keys_in[k:]), keys_out[k:]))
print("ku", ku[:50]); print("ki", ki[:50]); print("kj", kj[:50])
cu = {}; ci = {}; eku = []; eki = []; ekj = []
for t in range(k):
    eku.append(cu.get(tu[t], 0)); eki.append(ci.get(ti[t], 0)); ekj.append(ci.get(tj[t], 0))
    cu[tu[t]] = eku[-1] + 1; ci[ti[t]] = eki[-1] + 1; ci[tj[t]] = ekj[-1] + 1
print("expect ku", eku); print("expect ki", eki); print("expect kj", ekj)
n_users, n_items = len(g["users"]), len(g["items"])
cnt = ws[o_cnt:o_cnt + 4 * (n_users + n_items + 64)].view(np.int32)
print("cntU sum", cnt[:n_users].sum(), "cntI sum", cnt[n_users:n_users + n_items].sum(), "ticket", cnt[n_users + n_items])
