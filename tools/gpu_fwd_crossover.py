"""Forward solve (default Chain, 129 saves) for small batches: one column per CTA (solve_small_kernel) against the one-tile tensor-core kernel --
where should fwd_small_max sit?"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nfc_b200
syn = nfc_b200.synthetic
NN = nfc_b200.named_chain("dense-default", seed=0)
for p in NN.params(): p *= 1e-5
theta, _ = nfc_b200.destructure(NN)
sv = syn.saveat(129)
dall = syn.make_columns("ocean", B=65536, seed=2, K_CA=2.0)
for B in (100, 148, 222, 296, 444, 592, 888, 1184):
    T0, colp = dall["T0"][40000:40000 + B], dall["colp"][40000:40000 + B]
    out = []
    for name, small_max in (("one column per CTA", 10 ** 9), ("tensor cores", 0)):
        h = nfc_b200.Handle(NN.conv_width, NN.dense_spec, 1, sv); h.set_weights(theta)
        h.set_option("fwd_small_max", small_max)
        ts = []
        for i in range(4):
            r = h.solve(T0, colp, dall["glob"]); ts.append(h.counters()["last_kernel_ms"])
        out.append(f"{name}: {min(ts[1:]):.2f} ms")
        h.close()
    print(f"{B} columns | " + " | ".join(out), flush=True)
