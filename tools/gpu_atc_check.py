"""Tensor-core backward kernel (nfc_adjoint_tc.cuh) against the FP32 SIMT batch kernel and the C twin of the oracle.
Usage: python tools/gpu_atc_check.py [B ...]   (forces the batch kernels: NFC_ADJ_SMALL_MAX=0)"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["NFC_ADJ_SMALL_MAX"] = "0"
import nfc_b200
from oracle import c_oracle as CO
from oracle import nde_oracle as O
syn = nfc_b200.synthetic


def rel(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / np.linalg.norm(b))


def run(B, scale, K, n_save=33, oracle=True):
    d = syn.make_columns("training9") if B == 9 else syn.make_columns("ocean", B=B, seed=1)
    cw, layers = O.chain_spec("dense-default")
    theta = O.glorot_theta(cw, layers, 0, scale)
    sv = syn.saveat(n_save)
    colp = d["colp"].copy(); colp[:, 2] = K
    cp10 = colp.copy(); cp10[:, 2] = 10.0
    Ttrue = CO.solve(d["T0"], np.zeros_like(theta), cw, layers, cp10, d["glob"], sv, reltol=1e-6, abstol=1e-8, nthreads=8)["T"]
    out = {}
    for tag, env in (("simt", "0"), ("tc", "1")):
        os.environ["NFC_ADJ_TC"] = env
        kw = {}
        if os.environ.get("ATC_ADJ_RTOL"):
            kw = dict(adjoint_reltol=float(os.environ["ATC_ADJ_RTOL"]), adjoint_abstol=float(os.environ["ATC_ADJ_RTOL"]) * 1e-2)
        if os.environ.get("ATC_FIXED_DT"):
            kw = dict(fixed_dt=float(os.environ["ATC_FIXED_DT"]), adjoint_max_steps=int(1.0 / float(os.environ["ATC_FIXED_DT"])) + 8)
        h = nfc_b200.Handle(cw, layers, 1, sv, **kw)
        h.set_weights(theta)
        r = h.loss_grad(d["T0"], colp, d["glob"], Ttrue)
        out[tag] = (r, h.counters())
        del h
    (a, ca), (b, cb) = out["simt"], out["tc"]
    msg = f"B={B} scale={scale} K={K}: loss simt {a['loss']:.6e} tc {b['loss']:.6e} | grad rel(tc, simt) {rel(b['grad'], a['grad']):.2e} |g| {np.linalg.norm(a['grad']):.3e}"
    msg += f" | steps simt {ca['adj_steps_accepted']}+{ca['adj_steps_rejected']} tc {cb['adj_steps_accepted']}+{cb['adj_steps_rejected']}"
    msg += f" | status simt {np.unique(a['status']).tolist()} tc {np.unique(b['status']).tolist()} | ms simt {ca['last_adjoint_ms']:.1f} tc {cb['last_adjoint_ms']:.1f}"
    if oracle:
        okw = {}
        if os.environ.get("ATC_ADJ_RTOL"):
            okw = dict(adj_reltol=float(os.environ["ATC_ADJ_RTOL"]), adj_abstol=float(os.environ["ATC_ADJ_RTOL"]) * 1e-2)
        if os.environ.get("ATC_FIXED_DT"):
            okw = dict(fixed_dt=float(os.environ["ATC_FIXED_DT"]))
        o = CO.loss_grad_continuous(d["T0"], theta, cw, layers, colp, d["glob"], sv, Ttrue, nthreads=8, **okw)
        msg += f" | vs C twin: simt {rel(a['grad'], o['grad']):.2e} tc {rel(b['grad'], o['grad']):.2e}"
        # per-block errors (which accumulator is off, if any)
        N = [("W1", 0, 4096), ("b1", 4096, 4224), ("W2", 4224, 20608), ("b2", 20608, 20736), ("W3", 20736, 24704), ("b3", 24704, 24735)]
        msg += " | blocks " + " ".join(f"{n}:{rel(b['grad'][i:j], o['grad'][i:j]):.1e}" for n, i, j in N)
    print(msg, flush=True)


if __name__ == "__main__":
    Bs = [int(x) for x in sys.argv[1:]] or [9, 200]
    for B in Bs:
        for scale, K in ((1.0, 2.0), (1e-5, 2.0), (0.3, 0.0)):
            run(B, scale, K, oracle=B <= 512)
