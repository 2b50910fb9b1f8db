"""SURVEY row f3 measurement: compute_nde_solution_history-shaped workload (epochs x 21 simulations, 1153 saves) as ONE
nfc_solve_ensemble launch, next to (a) the same work as one nfc_solve per epoch on the GPU (what a line-by-line port of the
reference loop would do) and (b) the CPU oracle port on a few epochs."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["NFC_TC"] = "0"
import nfc_b200
from oracle import nde_oracle as O, c_oracle as CO
E = int(os.environ.get("NFC_ENS_E", "500")); S = 21; NS = 1153
syn = nfc_b200.synthetic
d = syn.make_columns("ocean", B=S, seed=4)
cw, layers = O.chain_spec()
sv = syn.saveat(NS)
base = O.glorot_theta(cw, layers, 0, 1.0)
rng = np.random.default_rng(0)
thetas = np.stack([(base * (1e-5 + 0.3 * e / E) + 0.01 * e / E * rng.standard_normal(base.size)).astype(np.float32) for e in range(E)])
h = nfc_b200.Handle(cw, layers, 1, sv)
T0 = np.ascontiguousarray(np.broadcast_to(d["T0"], (E, S, 32))); colp = np.ascontiguousarray(np.broadcast_to(d["colp"], (E, S, 3)))
for it in range(2):
    t0 = time.perf_counter(); r = h.solve_ensemble(thetas, T0, colp, d["glob"]); te = time.perf_counter() - t0
c = h.counters(); steps = c["steps_accepted"] + c["steps_rejected"]
print(f"ensemble: {E} members x {S} columns x {NS} saves: kernel {c['last_kernel_ms']:.1f} ms, call (host buffers) {te*1e3:.0f} ms, "
      f"{steps} column-steps ({steps/c['last_kernel_ms']*1e3:.3e}/s), bad status {int((r['status']!=0).sum())}")
n_ser = min(E, 40)
t0 = time.perf_counter()
for e in range(n_ser):
    h.set_weights(thetas[e]); one = h.solve(d["T0"], d["colp"], d["glob"])
ts = (time.perf_counter() - t0) / n_ser
assert np.array_equal(one["T"], r["T"][n_ser - 1])
print(f"serial GPU loop (set_weights + nfc_solve of {S} columns per epoch): {ts*1e3:.1f} ms per epoch -> {ts*E:.1f} s for {E} epochs; "
      f"ensemble launch is {ts*E/te:.0f}x faster end to end")
n_cpu = 4
t0 = time.perf_counter()
for e in range(n_cpu):
    cc = CO.solve(d["T0"], thetas[e * (E // n_cpu)], cw, layers, d["colp"], d["glob"], sv, nthreads=os.cpu_count())
tc = (time.perf_counter() - t0) / n_cpu
print(f"CPU oracle port ({os.cpu_count()} threads over the {S} columns): {tc*1e3:.0f} ms per epoch -> {tc*E:.0f} s for {E} epochs; ensemble launch is {tc*E/te:.0f}x faster")
