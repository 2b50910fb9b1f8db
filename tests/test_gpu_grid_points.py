"""--grid-points other than 32 (train_free_convection_nde.jl:29,90): the five Chains built on Nz < 32 levels, through the C ABI's
host-buffer calls, against the CPU oracle (generic in Nz; parity unpinned -- SURVEY.md 8c).  The library embeds such a Chain in its
32-level kernels (zero-padded weights, top flux on face Nz, norms over Nz levels): nothing of that is visible to the caller -- shapes are
[.., Nz] / [.., Nz + 1] and the parameter vector is Flux.destructure of the Nz-level Chain."""
import numpy as np
import pytest

from conftest import rel_l2
from oracle import nde_oracle as O

pytestmark = pytest.mark.gpu
RHS_TOL, TRAJ_TOL = 1e-5, 1e-3


def _setup(nfc, arch, Nz, B, n_save=9, seed=3, scale=1.0, **kw):
    d = nfc.synthetic.make_columns("ocean", B=B, seed=seed, Nz=Nz)
    cw, layers = O.chain_spec(arch, Nz)
    sv = nfc.synthetic.saveat(n_save)
    h = nfc.Handle(cw, layers, 1, sv, Nz=Nz, **kw)
    theta = O.glorot_theta(cw, layers, 4, scale)
    assert h.n_params == theta.size == O.n_params(cw, layers)
    h.set_weights(theta)
    return d, cw, layers, sv, h, theta


@pytest.mark.parametrize("arch,Nz", [("dense-default", 16), ("dense-default", 24), ("dense-default", 7), ("dense-wider", 16), ("dense-deeper", 12),
                                     ("conv-2", 16), ("conv-4", 20)])
def test_rhs_and_flux_vs_oracle(nfc, gpu_ok, arch, Nz):
    B = 203
    d, cw, layers, sv, h, theta = _setup(nfc, arch, Nz, B)
    rng = np.random.default_rng(Nz)
    T = (d["T0"] + 0.3 * rng.standard_normal(d["T0"].shape)).astype(np.float32)
    colp = d["colp"].copy()
    colp[:, 2] = rng.uniform(0, 10, B)
    f = h.rhs(T, colp, d["glob"])
    assert f.shape == (B, Nz)
    ref = O.rhs(T, theta, cw, layers, colp, d["glob"], np.float64)
    assert rel_l2(f, ref).max() < RHS_TOL
    pref = O.prefactor(d["glob"], np.float64)              # heat budget: the column mean moves by the boundary fluxes only
    np.testing.assert_allclose(f.astype(np.float64).sum(1) / Nz, -pref * (colp[:, 1].astype(np.float64) - colp[:, 0]), rtol=2e-3, atol=1e-4)
    Tprof = np.repeat(T[:, None, :], sv.size, axis=1)
    wT = h.diagnose_flux(Tprof, colp)
    assert wT.shape == (B, sv.size, Nz + 1)
    assert rel_l2(wT[:, 0], O.total_flux(T, theta, cw, layers, colp, np.float64)).max() < RHS_TOL


@pytest.mark.parametrize("arch,Nz", [("dense-default", 16), ("conv-2", 24), ("dense-wider", 12)])
def test_solve_vs_oracle(nfc, gpu_ok, arch, Nz):
    """Same float32 algorithm (controller-faithful Tsit5) on CPU and GPU, and the converged truth."""
    B = 40
    d, cw, layers, sv, h, theta = _setup(nfc, arch, Nz, B, scale=0.3)
    r = h.solve(d["T0"], d["colp"], d["glob"])
    assert r["T"].shape == (B, sv.size, Nz) and (r["status"] == 0).all()
    c = O.tsit5_solve(d["T0"], theta, cw, layers, d["colp"], d["glob"], sv, dtype=np.float32)
    assert rel_l2(r["T"], c["T"]).max() < TRAJ_TOL
    tot_g, tot_c = int((r["naccept"] + r["nreject"]).sum()), int((c["naccept"] + c["nreject"]).sum())
    assert abs(tot_g - tot_c) <= 0.05 * tot_c
    truth = O.truth_solve(d["T0"][:6], theta, cw, layers, d["colp"][:6], d["glob"], sv)
    assert rel_l2(r["T"][:6], truth).max() < TRAJ_TOL


@pytest.mark.parametrize("arch,Nz,K", [("dense-default", 16, 2.0), ("dense-default", 16, 0.0), ("conv-2", 12, 2.0)])
def test_loss_grad_vs_float64_autograd(nfc, gpu_ok, arch, Nz, K):
    """Tight tolerances (the kernel's own accuracy, not the step-sequence band): loss and gradient against float64 autograd on a fine grid."""
    B = 4
    d, cw, layers, sv, h, theta = _setup(nfc, arch, Nz, B, reltol=1e-6, abstol=1e-8)
    colp = d["colp"].copy()
    colp[:, 2] = K
    cp10 = colp.copy()
    cp10[:, 2] = 10.0
    Ttrue = O.tsit5_solve(d["T0"], np.zeros_like(theta), cw, layers, cp10, d["glob"], sv, 1e-8, 1e-10, np.float64)["T"].astype(np.float32)
    r = h.loss_grad(d["T0"], colp, d["glob"], Ttrue)
    assert (r["status"] == 0).all() and r["grad"].shape == theta.shape
    L, g, _ = O.loss_and_grad_autograd(d["T0"], theta, cw, layers, colp, d["glob"], sv, Ttrue, fixed_dt=1 / 1024)
    assert abs(r["loss"] - L) / L < 1e-3
    assert np.linalg.norm(r["grad"] - g) / np.linalg.norm(g) < 2e-3
    # the loss of loss_grad is the mse of the forward solve
    s = h.solve(d["T0"], colp, d["glob"])
    assert abs(O.mse(s["T"], Ttrue) - r["loss"]) / r["loss"] < 1e-4


def test_unsupported_combinations_fail_loudly(nfc, gpu_ok):
    cw, layers = O.chain_spec("dense-default", 16)
    sv = nfc.synthetic.saveat(5)
    with pytest.raises(nfc.NfcError):
        nfc.Handle(cw, layers, 1, sv, Nz=16, alg=1)                 # the stabilised integrator exists for 32 levels only
    with pytest.raises(nfc.NfcError):
        nfc.Handle(cw, layers, 1, sv, Nz=40)                        # more levels than a warp has lanes
    with pytest.raises(nfc.NfcError):
        nfc.Handle(cw, layers, 1, sv, Nz=20)                        # Chain built for another grid
    h = nfc.Handle(cw, layers, 1, sv, Nz=16)
    with pytest.raises(nfc.NfcError):
        h.set_option("tc", 1)


def test_training_loop_on_a_16_level_grid(nfc, gpu_ok):
    """train_neural_differential_equation (src/training.jl:34-78) with --grid-points 16: datasets, Chain and parameter vector live on 16
    levels; the host mirror passes Nz through to nfc_create."""
    syn = nfc.synthetic
    Nz, B, Nt = 16, 5, 65
    Ts, wTs = syn.reference_scalings(Nz)
    d = syn.make_columns("ocean", B=B, seed=7, Nz=Nz)
    cw, layers = O.chain_spec("dense-default", Nz)
    cp10 = d["colp"].copy()
    cp10[:, 2] = 10.0
    sv = syn.saveat(Nt)
    truth = O.tsit5_solve(d["T0"], np.zeros(O.n_params(cw, layers)), cw, layers, cp10, d["glob"], sv, 1e-6, 1e-8, np.float64)["T"]      # [B, Nt, Nz]
    zf = -256.0 + np.arange(Nz + 1) * (256.0 / Nz)
    datasets, params = {}, {}
    for i in range(B):
        flux = float(wTs.inv(d["colp"][i, 1]))
        datasets[i + 1] = nfc.ColumnDataset(T=Ts.inv(truth[i]).T, wT=np.zeros((Nz + 1, Nt)), zc=syn.cell_centres(Nz), zf=zf,
                                            times=np.arange(Nt) * 600.0 * 1152 / (Nt - 1), metadata={"temperature_flux": flux})
        params[i + 1] = nfc.ConvectiveAdjustmentNDEParameters(datasets[i + 1], Ts, wTs, 2)
    NN = nfc.named_chain("dense-default", Nz=Nz, seed=0)
    for p in NN.params():
        p *= 1e-5
    NN2 = nfc.train_neural_differential_equation(NN, nfc.ConvectiveAdjustmentNDE, params, nfc.Tsit5(), datasets, Ts, range(1, Nt + 1, 4), nfc.ADAM(), 4)
    losses = [h["mean_loss"] for h in NN2.history]
    assert len(losses) == 4 and np.isfinite(losses).all() and losses[-1] < losses[0]
    theta2, _ = nfc.destructure(NN2)
    assert theta2.size == O.n_params(cw, layers)
    # solve_nde with the trained Chain returns Nz x n_save
    nde = nfc.ConvectiveAdjustmentNDE(NN2, datasets[1], iterations=range(1, Nt + 1, 4))
    sol = nfc.solve_nde(nde, NN2, Ts(datasets[1].T[:, 0]).astype(np.float32), nfc.Tsit5(), params[1])
    assert np.asarray(sol).shape == (Nz, 17) and np.isfinite(sol).all()
