import os, sys, time, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nfc_b200
from oracle import c_oracle as CO
syn = nfc_b200.synthetic
d = syn.make_columns("training9")
NN = nfc_b200.named_chain("dense-default", seed=0)
for p in NN.params(): p *= 1e-5
theta, _ = nfc_b200.destructure(NN)
sv = syn.saveat(129)
Ttrue = np.repeat(d["T0"][:, None, :], 129, axis=1) * np.float32(0.98)
for opts in ({}, {"tc_record": 0}, {"tc": 0}):
    h = nfc_b200.Handle(NN.conv_width, NN.dense_spec, 1, sv); h.set_weights(theta)
    for k, v in opts.items(): h.set_option(k, v)
    for i in range(3):
        t0 = time.perf_counter(); r = h.loss_grad(d["T0"], d["colp"], d["glob"], Ttrue); t1 = time.perf_counter()
    c = h.counters()
    print(opts, "wall %.2f ms, record pass %.2f ms, total device %.2f ms, fwd steps %d adj steps %d" % ((t1 - t0) * 1e3, c["last_kernel_ms"], c["last_adjoint_ms"], c["steps_accepted"] + c["steps_rejected"], c["adj_steps_accepted"] + c["adj_steps_rejected"]), flush=True)
    for i in range(3):
        t0 = time.perf_counter(); s = h.solve(d["T0"], d["colp"], d["glob"]); t1 = time.perf_counter()
    print("   solve wall %.2f ms kernel %.2f ms" % ((t1 - t0) * 1e3, h.counters()["last_kernel_ms"]), flush=True)
