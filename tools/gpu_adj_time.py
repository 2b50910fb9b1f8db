import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nfc_b200
syn = nfc_b200.synthetic
NN = nfc_b200.named_chain("dense-default", seed=0)
for p in NN.params(): p *= 1e-5
theta, _ = nfc_b200.destructure(NN)
B = int(os.environ.get("NFC_ADJ_B", "16384"))
d = syn.make_columns("ocean", B=B, seed=2)
sv = syn.saveat(129)
h = nfc_b200.Handle(NN.conv_width, NN.dense_spec, 1, sv)
h.set_weights(theta)
for tag, Ttrue in (("T0+0.01", np.repeat(d["T0"][:, None, :], 129, axis=1) + 0.01), ("0.98*T0", np.repeat(d["T0"][:, None, :], 129, axis=1) * np.float32(0.98))):
    for i in range(int(os.environ.get("NFC_ADJ_IT", "2"))):      # the second and later calls use the backward-step history as the queue key
        r = h.loss_grad(d["T0"], d["colp"], d["glob"], Ttrue)
        c = h.counters()
        print(f"   call {i}: total {c['last_adjoint_ms']:.1f} ms", flush=True)
    a = h.adjoint_steps(B); fw = r["status"]
    print("   adj steps/column: min %d p50 %d p90 %d p99 %d max %d mean %.0f" % (a.min(), np.percentile(a,50), np.percentile(a,90), np.percentile(a,99), a.max(), a.mean()), flush=True)
    print(f"{tag}: fwd(record) {c['last_kernel_ms']:.1f} ms total {c['last_adjoint_ms']:.1f} ms, fwd steps {c['steps_accepted']+c['steps_rejected']}, adj steps {c['adj_steps_accepted']}+{c['adj_steps_rejected']}, loss {r['loss']:.6e}", flush=True)
