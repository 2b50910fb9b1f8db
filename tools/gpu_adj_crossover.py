"""Backward pass of nfc_loss_grad (default Chain, 129 saves): one column per CTA against the tensor-core kernel (+ side launch of the longest
columns from the second call on) for mid-size batches -- where should adj_small_max sit?"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nfc_b200
syn = nfc_b200.synthetic
NN = nfc_b200.named_chain("dense-default", seed=0)
for p in NN.params(): p *= 1e-5
theta, _ = nfc_b200.destructure(NN)
sv = syn.saveat(129)
dall = syn.make_columns("ocean", B=65536, seed=2)
for B in (500, 1000, 1500, 2000, 3000, 4000, 7000):
    T0, colp = dall["T0"][40000:40000 + B], dall["colp"][40000:40000 + B]
    Ttrue = np.repeat(T0[:, None, :], 129, axis=1) * np.float32(0.98)
    out = []
    for name, small_max, heavy in (("one column per CTA", 10 ** 9, 0), ("tensor cores", 0, 0), ("tensor cores + side launch", 0, max(8, min(48, B // 32)))):
        h = nfc_b200.Handle(NN.conv_width, NN.dense_spec, 1, sv); h.set_weights(theta)
        h.set_option("adj_small_max", small_max); h.set_option("adj_heavy", heavy)
        ts = []
        for i in range(3):
            r = h.loss_grad(T0, colp, dall["glob"], Ttrue); c = h.counters(); ts.append(c["last_adjoint_ms"] - c["last_kernel_ms"])
        out.append(f"{name}: first {ts[0]:.1f} later {min(ts[1:]):.1f} ms")
        h.close()
    print(f"{B} columns | " + " | ".join(out), flush=True)
