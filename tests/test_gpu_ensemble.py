"""GPU parity tests of SURVEY row f3, the epoch-history evaluator (src/testing.jl:42-79): nfc_solve_ensemble integrates
every (member = epoch's Chain, simulation) pair in one launch.  Checked bit for bit against one nfc_solve per member with
the FP32 kernels, against the CPU oracle, and through the reference-shaped compute_nde_solution_history."""
import numpy as np
import pytest

from conftest import rel_l2
from oracle import c_oracle as CO
from oracle import nde_oracle as O

pytestmark = pytest.mark.gpu
TRAJ_TOL = 1e-3


def _members(cw, layers, E, seed0=0):
    # a training history in miniature: member e = Glorot weights of seed e, scaled differently
    return np.stack([O.glorot_theta(cw, layers, seed0 + e, 0.05 + 0.1 * (e % 7)) for e in range(E)]).astype(np.float32)


@pytest.mark.parametrize("arch,E,S", [("dense-default", 5, 21), ("dense-default", 3, 70), ("dense-default", 301, 3),
                                      ("dense-wider", 4, 9), ("dense-deeper", 4, 21), ("conv-2", 4, 21), ("conv-4", 2, 1)])
def test_ensemble_equals_one_solve_per_member(nfc, gpu_ok, monkeypatch, arch, E, S):
    """Members, slots and queue order must not leak into the arithmetic: every member's columns are bit-identical to a
    plain nfc_solve with that member's weights (FP32 kernels).  S = 70 exercises slot refill inside a member, E = 301 the
    member queue (more members than SMs), S = 1 nearly empty warps."""
    monkeypatch.setenv("NFC_TC", "0")
    cw, layers = O.chain_spec(arch)
    sv = nfc.synthetic.saveat(9)
    d = nfc.synthetic.make_columns("ocean", B=S, seed=3)
    thetas = _members(cw, layers, E)
    h = nfc.Handle(cw, layers, 1, sv)
    T0 = np.broadcast_to(d["T0"], (E, S, 32))
    colp = np.broadcast_to(d["colp"], (E, S, 3))
    r = h.solve_ensemble(thetas, T0, colp, d["glob"])
    assert r["T"].shape == (E, S, 9, 32) and (r["status"] == 0).all()
    cnt = h.counters()
    assert cnt["columns"] == E * S and cnt["steps_accepted"] == int(r["naccept"].sum())
    for e in ([0, E // 2, E - 1] if E > 8 else range(E)):
        h.set_weights(thetas[e])
        one = h.solve(d["T0"], d["colp"], d["glob"])
        assert np.array_equal(one["T"], r["T"][e]), f"member {e}"
        assert np.array_equal(one["naccept"], r["naccept"][e]) and np.array_equal(one["nreject"], r["nreject"][e])


def test_ensemble_members_differ_and_match_oracle(nfc, gpu_ok):
    cw, layers = O.chain_spec()
    sv = nfc.synthetic.saveat(17)
    d = nfc.synthetic.make_columns("training9")
    S = d["T0"].shape[0]
    thetas = _members(cw, layers, 3, seed0=10)
    h = nfc.Handle(cw, layers, 1, sv)
    r = h.solve_ensemble(thetas, np.broadcast_to(d["T0"], (3, S, 32)), np.broadcast_to(d["colp"], (3, S, 3)), d["glob"])
    assert not np.allclose(r["T"][0], r["T"][1])
    for e in range(3):
        c = CO.solve(d["T0"], thetas[e], cw, layers, d["colp"], d["glob"], sv, nthreads=8)
        assert rel_l2(r["T"][e], c["T"]).max() < TRAJ_TOL


def test_ensemble_argument_errors(nfc, gpu_ok):
    cw, layers = O.chain_spec()
    h = nfc.Handle(cw, layers, 1, nfc.synthetic.saveat(5))
    d = nfc.synthetic.make_columns("training9")
    with pytest.raises(ValueError):
        h.solve_ensemble(np.zeros((2, 7), np.float32), np.zeros((2, 9, 32)), np.zeros((2, 9, 3)), d["glob"])
    with pytest.raises(ValueError):
        h.solve_ensemble(np.zeros((2, h.n_params), np.float32), np.zeros((3, 9, 32)), np.zeros((3, 9, 3)), d["glob"])
    r = h.solve_ensemble(np.zeros((0, h.n_params), np.float32), np.zeros((0, 9, 32)), np.zeros((0, 9, 3)), d["glob"])
    assert r["T"].shape == (0, 9, 5, 32)


def test_compute_nde_solution_history_matches_serial_solve_nde(nfc, gpu_ok, tmp_path, monkeypatch):
    """The reference loop (src/testing.jl:62-73): for each epoch, for each dataset, solve_nde(ds, NN_e, ...) and the per-time
    mse against ds["T"].  The one-launch history must reproduce that loop."""
    monkeypatch.setenv("NFC_TC", "0")
    syn = nfc.synthetic
    Ts, wTs = syn.reference_scalings()
    d = syn.make_columns("training9")
    Nt = 33
    zf = -256.0 + np.arange(33) * 8.0
    datasets, params = {}, {}
    for i in range(3):
        T = np.repeat(Ts.inv(d["T0"][i])[:, None], Nt, axis=1) + 1e-3 * np.arange(Nt)[None, :]
        datasets[i + 1] = nfc.ColumnDataset(T=T, wT=np.zeros((33, Nt)), zc=syn.cell_centres(), zf=zf, times=np.arange(Nt) * 21600.0,
                                            metadata={"temperature_flux": float(syn.heat_flux((1 + i) * 1e-8))})
        params[i + 1] = nfc.ConvectiveAdjustmentNDEParameters(datasets[i + 1], Ts, wTs, 2)
    NN = nfc.named_chain("dense-default", seed=0)
    theta0, reconstruct = nfc.destructure(NN)
    hist = {f"neural_network/{e}": (theta0 * (0.1 * e)).astype(np.float32) for e in range(1, 5)}
    path = str(tmp_path / "history.npz")
    np.savez(path, **hist)
    sol, loss = nfc.compute_nde_solution_history(datasets, nfc.ConvectiveAdjustmentNDE, params, nfc.Tsit5(), NN, path, Ts)
    assert set(sol) == {1, 2, 3} and sol[1].shape == (4, 32, Nt) and loss[1].shape == (4, Nt)
    for e in (1, 4):
        NNe = reconstruct(hist[f"neural_network/{e}"])
        for i in (1, 3):
            ref = nfc.solve_nde(datasets[i], NNe, nfc.ConvectiveAdjustmentNDE, params[i], nfc.Tsit5(), Ts, wTs)["T"]
            np.testing.assert_allclose(sol[i][e - 1], ref, rtol=0, atol=1e-5 * np.abs(ref).max())
            np.testing.assert_allclose(loss[i][e - 1], np.mean((datasets[i].T - ref) ** 2, axis=0), rtol=1e-3, atol=1e-12)
