import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nfc_b200
syn = nfc_b200.synthetic
d = syn.make_columns("ocean", B=1048576, seed=2, K_CA=2.0)
NN = nfc_b200.named_chain("dense-default", seed=0)
for p in NN.params(): p *= 1e-5
theta, _ = nfc_b200.destructure(NN)
sv = syn.saveat(129)
idx = [426987, 761724]
T0 = d["T0"][idx]; colp = d["colp"][idx]
Tt = np.repeat(T0[:, None, :], 129, axis=1) * np.float32(0.98)
def run(tag, env={}, **kw):
    for k, v in env.items(): os.environ[k] = v
    h = nfc_b200.Handle(NN.conv_width, NN.dense_spec, 1, sv, max_iters=20000, **kw); h.set_weights(theta)
    t0 = time.perf_counter(); r = h.loss_grad(T0, colp, d["glob"], Tt); dt = time.perf_counter() - t0
    c = h.counters()
    print(f"{tag}: {dt*1e3:.0f} ms status {r['status'].tolist()} adj steps {h.adjoint_steps(2).tolist()} acc {c['adj_steps_accepted']} rej {c['adj_steps_rejected']} fwd {c['steps_accepted']}+{c['steps_rejected']} loss {r['loss']:.6e}", flush=True)
    for k in env: del os.environ[k]
run("default (small kernel, tc record)")
run("simt record", {"NFC_TC_RECORD": "0"})
run("tc off entirely", {"NFC_TC": "0"})
run("adjoint reltol 2e-4", adjoint_reltol=2e-4)
run("adjoint reltol 5e-5", adjoint_reltol=5e-5)
run("reltol 1e-5 fwd", reltol=1e-5, adjoint_reltol=1e-4)
tapes = {}
for tag, env in (("tc", {}), ("simt", {"NFC_TC_RECORD": "0"})):
    for k, v in env.items(): os.environ[k] = v
    h = nfc_b200.Handle(NN.conv_width, NN.dense_spec, 1, sv, max_iters=3000); h.set_weights(theta)
    r = h.loss_grad(T0, colp, d["glob"], Tt)
    tapes[tag] = h.tape(0)
    for k in env: del os.environ[k]
a, b = tapes["tc"], tapes["simt"]
print("rows", a.shape, b.shape, "finite", np.isfinite(a).all(), np.isfinite(b).all())
print("t tc  ", a[:, 160][:12], "...", a[:, 160][-6:])
print("h tc  ", a[:, 161][:12], "...", a[:, 161][-6:])
print("t+h - next t (tc):  max abs", np.abs(a[:-1, 160] + a[:-1, 161] - a[1:, 160]).max(), " end", a[-1, 160] + a[-1, 161])
print("t+h - next t (simt):max abs", np.abs(b[:-1, 160] + b[:-1, 161] - b[1:, 160]).max(), " end", b[-1, 160] + b[-1, 161])
n = min(len(a), len(b))
for k, nm in enumerate(["u", "c1", "c2", "c3", "c4"]):
    da = np.abs(a[:n, 32 * k:32 * k + 32] - b[:n, 32 * k:32 * k + 32]).max(axis=1)
    print(nm, "max |tc - simt| per row (first 10 / worst):", da[:10], da.max(), int(da.argmax()), " scale", np.abs(b[:n, 32 * k:32 * k + 32]).max())
np.savez("gpurun_out/badcol_tapes.npz", tc=a, simt=b)
