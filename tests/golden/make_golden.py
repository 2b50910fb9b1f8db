"""Generates tests/golden/*.npz from the float64 NumPy oracle (oracle/nde_oracle.py) and scipy DOP853 truth.
The reference is Julia and cannot run here: these vectors pin the restatement, not the Julia code.
Run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import nfc_b200  # noqa: E402
from oracle import nde_oracle as O  # noqa: E402

syn = nfc_b200.synthetic
OUT = os.path.dirname(os.path.abspath(__file__))


def main():
    d = syn.make_columns("training9")
    rng = np.random.Generator(np.random.PCG64(7))
    # ---- RHS golden: all five Chains, unscaled Glorot weights, perturbed profiles, K_CA in {0, 2} ----
    for arch in ["dense-default", "dense-wider", "dense-deeper", "conv-2", "conv-4"]:
        cw, layers = O.chain_spec(arch)
        theta = O.glorot_theta(cw, layers, seed=0, scale=1.0)
        theta[-31:] = rng.uniform(-0.1, 0.1, 31).astype(np.float32)          # non-zero output bias
        T = (d["T0"] + 0.25 * rng.standard_normal(d["T0"].shape)).astype(np.float32)
        colp = d["colp"].copy()
        colp[::2, 2] = 0.0
        f = O.rhs(T, theta, cw, layers, colp, d["glob"], np.float64)
        np.savez_compressed(os.path.join(OUT, f"rhs_{arch}.npz"), theta=theta, T=T, colp=colp, glob=d["glob"], dTdt=f)
    # ---- trajectory golden: default Chain, 3 weight scalings x K_CA in {0, 2}, 33 save points ----
    cw, layers = O.chain_spec("dense-default")
    sv = syn.saveat(33)
    for tag, scale in [("1e-5", 1e-5), ("0p3", 0.3), ("1p0", 1.0)]:
        theta = O.glorot_theta(cw, layers, seed=0, scale=scale)
        for K in (0.0, 2.0):
            colp = d["colp"].copy()
            colp[:, 2] = K
            t64 = O.tsit5_solve(d["T0"], theta, cw, layers, colp, d["glob"], sv, dtype=np.float64)
            truth = O.truth_solve(d["T0"], theta, cw, layers, colp, d["glob"], sv, rtol=1e-11, atol=1e-11)
            np.savez_compressed(os.path.join(OUT, f"traj_{tag}_K{int(K)}.npz"), theta_seed=0, theta_scale=scale,
                                T0=d["T0"], colp=colp, glob=d["glob"], saveat=sv, T_tsit5_f64=t64["T"].astype(np.float32),
                                T_truth=truth.astype(np.float32), naccept=t64["naccept"], nreject=t64["nreject"])
    # ---- gradient golden: float64 autograd, fixed step (well-conditioned), K_CA = 0 and 2, 17 saves, 3 columns ----
    sv = syn.saveat(17)
    theta = O.glorot_theta(cw, layers, seed=0, scale=1.0)
    cp10 = d["colp"].copy()
    cp10[:, 2] = 10.0
    Ttrue = O.tsit5_solve(d["T0"], np.zeros_like(theta), cw, layers, cp10, d["glob"], sv, dtype=np.float64, reltol=1e-9,
                          abstol=1e-11)["T"].astype(np.float32)
    for K in (0.0, 2.0):
        colp = d["colp"][:3].copy()
        colp[:, 2] = K
        L, g, dT0 = O.loss_and_grad_autograd(d["T0"][:3], theta, cw, layers, colp, d["glob"], sv, Ttrue[:3], fixed_dt=1 / 512)
        np.savez_compressed(os.path.join(OUT, f"grad_K{int(K)}.npz"), theta_seed=0, theta_scale=1.0, T0=d["T0"][:3],
                            colp=colp, glob=d["glob"], saveat=sv, Ttrue=Ttrue[:3], fixed_dt=1 / 512, loss=L,
                            grad=g.astype(np.float32), dT0=dT0.astype(np.float32))
    print("golden written to", OUT)


if __name__ == "__main__":
    main()
