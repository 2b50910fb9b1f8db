"""Which columns of the config-4 set fail in the backward pass, and do they fail in every backward kernel?"""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nfc_b200
syn = nfc_b200.synthetic
B = 1048576
d = syn.make_columns("ocean", B=B, seed=2, K_CA=2.0)
NN = nfc_b200.named_chain("dense-default", seed=0)
for p in NN.params(): p *= 1e-5
theta, _ = nfc_b200.destructure(NN)
sv = syn.saveat(129)
dev = torch.device("cuda", 0)
h = nfc_b200.Handle(NN.conv_width, NN.dense_spec, 1, sv)
h.set_weights(theta)
bad = []
CH = 131072
for k in range(B // CH):
    lo, hi = k * CH, (k + 1) * CH
    T0 = torch.from_numpy(d["T0"][lo:hi]).to(dev); colp = torch.from_numpy(d["colp"][lo:hi]).to(dev)
    Tt = (T0 * 0.98).unsqueeze(1).expand(CH, 129, 32).contiguous()
    g = torch.empty(h.n_params, device=dev); ls = torch.zeros(2, dtype=torch.float64, device=dev); st = torch.zeros(CH, dtype=torch.int32, device=dev)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    h.loss_grad_dev(T0, colp, d["glob"], Tt, ls, g, st, n_total_columns=CH)
    h.sync(); torch.cuda.synchronize(); dt = time.perf_counter() - t0
    a = h.adjoint_steps(CH)
    s = st.cpu().numpy()
    idx = np.nonzero(s)[0]
    print(f"chunk {k}: {dt*1e3:.0f} ms, adj steps max {a.max()} p99.9 {np.percentile(a, 99.9):.0f}, bad {[(int(i)+lo, int(s[i]), int(a[i])) for i in idx]}", flush=True)
    bad += [int(i) + lo for i in idx]
    top = np.argsort(a)[-3:]
    print("   longest:", [(int(i) + lo, int(a[i])) for i in top], flush=True)
    bad += [int(i) + lo for i in top[-1:]]
bad = sorted(set(bad))
print("columns to re-run alone:", bad)
for tag, env in (("small kernel", {}), ("simt batch", {"NFC_ADJ_SMALL_MAX": "0", "NFC_ADJ_TC": "0"}), ("tc batch", {"NFC_ADJ_SMALL_MAX": "0", "NFC_ADJ_TC": "1"})):
    for k_, v_ in env.items(): os.environ[k_] = v_
    h2 = nfc_b200.Handle(NN.conv_width, NN.dense_spec, 1, sv); h2.set_weights(theta)
    T0 = d["T0"][bad]; colp = d["colp"][bad]
    Tt = np.repeat(T0[:, None, :], 129, axis=1) * np.float32(0.98)
    t0 = time.perf_counter()
    r = h2.loss_grad(T0, colp, d["glob"], Tt)
    dt = time.perf_counter() - t0
    print(tag, f"{dt*1e3:.0f} ms", "status", r["status"].tolist(), "adj steps", h2.adjoint_steps(len(bad)).tolist(), "|g|", float(np.linalg.norm(r["grad"])), flush=True)
    for k_ in env: del os.environ[k_]
print("colp of those columns:", d["colp"][bad].tolist())
