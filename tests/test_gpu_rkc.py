"""GPU parity tests of SURVEY row f2: the stabilised explicit integrator (RKC2, in the role of ROCK4) behind nfc_solve with
alg = NFC_ALG_RKC2, against its CPU twin oracle/nde_oracle.py:rkc_solve (same float32 algorithm, same coefficient table),
against the high-accuracy truth integrator, and against the Tsit5 path (same ODE, so the same trajectories to tolerance)."""
import numpy as np
import pytest

from conftest import rel_l2
from oracle import nde_oracle as O

pytestmark = pytest.mark.gpu
TRAJ_TOL = 1e-3


def _cols(nfc, B, seed, K=None):
    d = nfc.synthetic.make_columns("ocean", B=B, seed=seed)
    if K is not None:
        d["colp"][:, 2] = K
    return d


@pytest.mark.parametrize("arch,tc", [("dense-default", "1"), ("dense-default", "0"), ("dense-default", "bf16"), ("dense-wider", "0"),
                                     ("dense-deeper", "0"), ("conv-2", "0"), ("conv-4", "0")])
def test_rkc_matches_cpu_twin_all_chains(nfc, gpu_ok, monkeypatch, arch, tc):
    """tc = "1": tensor-core kernel (FP16x2 split, the default for the default Chain), "bf16": its BF16x3 split, "0": FP32 kernel."""
    monkeypatch.setenv("NFC_TC", "0" if tc == "0" else "1")
    monkeypatch.setenv("NFC_TC_SPLIT", "bf16" if tc == "bf16" else "fp16")
    B = 19 if tc == "0" else 150
    d = _cols(nfc, B, 21)
    sv = nfc.synthetic.saveat(17)
    cw, layers = O.chain_spec(arch)
    theta = O.glorot_theta(cw, layers, 2, 0.3)
    h = nfc.Handle(cw, layers, 1, sv, alg="RKC2")
    h.set_weights(theta)
    r = h.solve(d["T0"], d["colp"], d["glob"])
    c = O.rkc_solve(d["T0"][:19], theta, cw, layers, d["colp"][:19], d["glob"], sv, dtype=np.float32)
    assert (r["status"] == 0).all() and (c["status"] == 0).all()
    full = r
    r = {k: v[:19] for k, v in r.items()}
    assert rel_l2(r["T"], c["T"]).max() < TRAJ_TOL
    tot_g, tot_c = int((r["naccept"] + r["nreject"]).sum()), int((c["naccept"] + c["nreject"]).sum())
    assert abs(tot_g - tot_c) <= 0.1 * tot_c
    cnt = h.counters()
    assert cnt["steps_accepted"] == int(full["naccept"].sum()) and cnt["steps_rejected"] == int(full["nreject"].sum())
    assert abs(cnt["rhs_evals"] * 19 / B - int(c["nfe"].sum())) <= 0.1 * int(c["nfe"].sum())


@pytest.mark.parametrize("K", [0.0, 2.0, 10.0])
def test_rkc_converges_to_the_truth_and_agrees_with_tsit5(nfc, gpu_ok, K):
    d = nfc.synthetic.make_columns("training9")
    d["colp"][:, 2] = K
    sv = nfc.synthetic.saveat(33)
    cw, layers = O.chain_spec()
    theta = O.glorot_theta(cw, layers, 0, 0.3)
    hr = nfc.Handle(cw, layers, 1, sv, alg="RKC2")
    ht = nfc.Handle(cw, layers, 1, sv)
    hr.set_weights(theta); ht.set_weights(theta)
    r = hr.solve(d["T0"], d["colp"], d["glob"])
    t = ht.solve(d["T0"], d["colp"], d["glob"])
    truth = O.truth_solve(d["T0"], theta, cw, layers, d["colp"], d["glob"], sv)
    assert (r["status"] == 0).all()
    assert rel_l2(r["T"], truth).max() < TRAJ_TOL
    assert rel_l2(r["T"], t["T"]).max() < TRAJ_TOL
    # heat budget: the column integral changes by -pref (top - bottom) t whatever the integrator does in between
    pref = d["glob"][1] / d["glob"][0] * d["glob"][3] / d["glob"][2]
    budget = r["T"].sum(axis=2) - r["T"][:, :1].sum(axis=2)
    want = -pref * 32 * (d["colp"][:, 1] - d["colp"][:, 0])[:, None] * sv[None, :]
    assert np.abs(budget - want).max() < 2e-3 * max(1.0, np.abs(want).max())


def test_rkc_needs_fewer_evaluations_when_adjustment_is_stiff(nfc, gpu_ok):
    """The point of the method (figureC's K_CA sweep uses ROCK4): at K_CA = 10 Tsit5 is stability-limited, RKC2 is not."""
    B = 512
    d = _cols(nfc, B, 8, K=10.0)
    sv = nfc.synthetic.saveat(129)
    cw, layers = O.chain_spec()
    theta = O.glorot_theta(cw, layers, 0, 0.3)
    hr = nfc.Handle(cw, layers, 1, sv, alg="ROCK4")          # the reference's name selects the same method
    ht = nfc.Handle(cw, layers, 1, sv)
    hr.set_weights(theta); ht.set_weights(theta)
    r = hr.solve(d["T0"], d["colp"], d["glob"]); er = hr.counters()["rhs_evals"]
    t = ht.solve(d["T0"], d["colp"], d["glob"]); et = ht.counters()["rhs_evals"]
    assert (r["status"] == 0).all() and (t["status"] == 0).all()
    assert rel_l2(r["T"], t["T"]).max() < TRAJ_TOL
    assert er < 0.5 * et, (er, et)


@pytest.mark.parametrize("tc", ["1", "0"])
def test_rkc_columns_are_independent_and_edge_cases(nfc, gpu_ok, monkeypatch, tc):
    monkeypatch.setenv("NFC_TC", tc)
    B = 333
    d = _cols(nfc, B, 13)
    sv = nfc.synthetic.saveat(5)
    cw, layers = O.chain_spec()
    h = nfc.Handle(cw, layers, 1, sv, alg="RKC2")
    h.set_weights(O.glorot_theta(cw, layers, 2, 0.3))
    r = h.solve(d["T0"], d["colp"], d["glob"])
    perm = np.random.default_rng(1).permutation(B)
    rp = h.solve(d["T0"][perm], d["colp"][perm], d["glob"])
    assert np.array_equal(rp["T"], r["T"][perm]) and np.array_equal(rp["naccept"], r["naccept"][perm])
    one = h.solve(d["T0"][7:8], d["colp"][7:8], d["glob"])
    assert np.array_equal(one["T"][0], r["T"][7])
    # zero network, no adjustment, no fluxes: f == 0, spectral radius 0, the profile must not move
    h.set_weights(np.zeros(h.n_params, np.float32))
    colp = np.zeros((4, 3), np.float32)
    z = h.solve(d["T0"][:4], colp, d["glob"])
    # (to rounding: the stage recurrence mu Y_{j-1} + nu Y_{j-2} + (1 - mu - nu) y_n reproduces a constant only to an ulp)
    assert (z["status"] == 0).all()
    np.testing.assert_allclose(z["T"], np.repeat(d["T0"][:4, None, :], 5, axis=1), rtol=2e-6, atol=1e-6)
    # zero network with fluxes: only the boundary cells drift, linearly (closed form)
    colp[:, 0], colp[:, 1] = 0.1, 0.3
    z = h.solve(d["T0"][:4], colp, d["glob"])
    pref = d["glob"][1] / d["glob"][0] * d["glob"][3] / d["glob"][2]
    np.testing.assert_allclose(z["T"][:, -1, 31], d["T0"][:4, 31] - pref * 0.3 * 32, rtol=2e-5, atol=1e-5)
    np.testing.assert_allclose(z["T"][:, -1, 0], d["T0"][:4, 0] + pref * 0.1 * 32, rtol=2e-5, atol=1e-5)
    np.testing.assert_allclose(z["T"][:, :, 1:31], np.repeat(d["T0"][:4, None, 1:31], 5, axis=1), rtol=2e-6, atol=1e-6)


def test_rkc_handle_rejects_what_it_cannot_do(nfc, gpu_ok):
    cw, layers = O.chain_spec()
    sv = nfc.synthetic.saveat(5)
    with pytest.raises(nfc.NfcError):
        nfc.Handle(cw, layers, 1, sv, alg="RKC2", fixed_dt=0.01)
    with pytest.raises(nfc.NfcError):
        nfc.Handle(cw, layers, 1, sv, alg=7)


def test_reference_api_selects_the_stabilised_method(nfc, gpu_ok):
    """solve_nde(nde, NN, T0, ROCK4(), nde_params) (train_free_convection_nde.jl:328 passes ROCK4 as `algorithm`)."""
    syn = nfc.synthetic
    Ts, wTs = syn.reference_scalings()
    d = syn.make_columns("single")
    zf = -256.0 + np.arange(33) * 8.0
    Tds = np.repeat(Ts.inv(d["T0"][0])[:, None], 1153, axis=1)
    ds = nfc.ColumnDataset(T=Tds, wT=np.zeros((33, 1153)), zc=syn.cell_centres(), zf=zf, times=np.arange(1153) * 600.0,
                           metadata={"temperature_flux": float(syn.heat_flux(1e-8))})
    NN = nfc.named_chain("dense-default", seed=0)
    for l in NN.layers:
        l.weight *= 0.3
    p = nfc.ConvectiveAdjustmentNDEParameters(ds, Ts, wTs, 2)
    nde = nfc.ConvectiveAdjustmentNDE(NN, ds, iterations=range(1, 1154, 9))
    a = nfc.solve_nde(nde, NN, Ts(ds.T[:, 0]), nfc.ROCK4(), p)
    b = nfc.solve_nde(nde, NN, Ts(ds.T[:, 0]), nfc.Tsit5(), p)
    assert a.shape == (32, 129)
    assert rel_l2(a.T[None], b.T[None]).max() < TRAJ_TOL
    # the 7-argument form (src/solve.jl:8-51) with the production integrator: the flux diagnosis must use the handle that holds the weights
    full = nfc.solve_nde(ds, NN, nfc.ConvectiveAdjustmentNDE, p, nfc.ROCK4(), Ts, wTs)
    ref = nfc.solve_nde(ds, NN, nfc.ConvectiveAdjustmentNDE, p, nfc.Tsit5(), Ts, wTs)
    assert full["T"].shape == (32, 1153) and full["wT"].shape == (33, 1153)
    assert rel_l2(full["T"].T[None], ref["T"].T[None]).max() < TRAJ_TOL
    assert np.isfinite(full["wT"]).all()


def test_loss_grad_through_the_stabilised_integrator(nfc, gpu_ok):
    """nfc_loss_grad on an RKC2 handle (the production configuration trains with ROCK4, train_free_convection_nde.sh:3): the forward pass
    records residuals and a cubic-Hermite tape, the continuous adjoint reads it like the Tsit5 tape.  Same continuous problem as with
    Tsit5, so loss and gradient agree to the integrators' tolerance; both against the C twin (Tsit5) of the oracle."""
    from oracle import c_oracle as CO
    syn = nfc.synthetic
    d = syn.make_columns("training9")
    cw, layers = O.chain_spec("dense-default")
    sv = syn.saveat(33)
    cp10 = d["colp"].copy(); cp10[:, 2] = 10.0
    for scale, K, tol in ((1.0, 2.0, 5e-3), (0.3, 0.0, 3e-3)):
        theta = O.glorot_theta(cw, layers, 0, scale)
        colp = d["colp"].copy(); colp[:, 2] = K
        Ttrue = CO.solve(d["T0"], np.zeros_like(theta), cw, layers, cp10, d["glob"], sv, reltol=1e-6, abstol=1e-8, nthreads=8)["T"]
        res = {}
        for alg in ("Tsit5", "RKC2"):
            h = nfc.Handle(cw, layers, 1, sv, alg=alg, reltol=1e-5, abstol=1e-7, adjoint_reltol=1e-5, adjoint_abstol=1e-7)
            h.set_weights(theta)
            res[alg] = h.loss_grad(d["T0"], colp, d["glob"], Ttrue)
            assert (res[alg]["status"] == 0).all()
        a, b = res["Tsit5"], res["RKC2"]
        assert abs(a["loss"] - b["loss"]) / a["loss"] < 1e-3
        rel = float(np.linalg.norm(a["grad"].astype(np.float64) - b["grad"]) / np.linalg.norm(a["grad"]))
        assert rel < tol, rel
    # default tolerances, larger batch (tensor-core RKC2 RECORD pass + batch backward kernel): finite, status clean, close to Tsit5
    d = syn.make_columns("ocean", B=300, seed=5)
    theta = O.glorot_theta(cw, layers, 0, 0.3)
    colp = d["colp"].copy(); colp[:, 2] = 2.0
    Ttrue = np.repeat(d["T0"][:, None, :], 33, axis=1) * np.float32(0.98)
    out = {}
    for alg in ("Tsit5", "RKC2"):
        h = nfc.Handle(cw, layers, 1, sv, alg=alg)
        h.set_weights(theta)
        out[alg] = h.loss_grad(d["T0"], colp, d["glob"], Ttrue)
        assert (out[alg]["status"] == 0).all() and np.isfinite(out[alg]["grad"]).all()
    assert abs(out["Tsit5"]["loss"] - out["RKC2"]["loss"]) / out["Tsit5"]["loss"] < 2e-3
    cos = float(out["Tsit5"]["grad"].astype(np.float64) @ out["RKC2"]["grad"] / np.linalg.norm(out["Tsit5"]["grad"]) / np.linalg.norm(out["RKC2"]["grad"]))
    assert cos > 0.999
