"""embedded path (explicit NN flux step + implicit convective adjustment as a per-column tridiagonal solve,
src/oceananigans_nn.jl:3-30,158-282).  CPU: the oracle restatement; GPU: nfc_embedded_solve vs the oracle."""
import numpy as np
import pytest

from conftest import rel_l2
from oracle import nde_oracle as O


def _columns(nfc, B, seed=0):
    syn = nfc.synthetic
    Ts, wTs = syn.reference_scalings()
    d = syn.make_columns("ocean", B=B, seed=seed)
    T0 = Ts.inv(d["T0"]).astype(np.float32)                         # physical temperatures
    Q = wTs.inv(d["colp"][:, 1]).astype(np.float32)                 # physical surface flux
    return T0, Q, [Ts.mu, Ts.sigma, wTs.mu, wTs.sigma]


def test_tridiagonal_matches_dense_solve():
    rng = np.random.default_rng(0)
    T = np.cumsum(rng.standard_normal(32)) * 0.01 + 20
    ld, d, ud = O.convective_adjustment_matrix(T, 600.0, 8.0, 10.0, 1e-3)
    A = np.diag(d) + np.diag(ld, -1) + np.diag(ud, 1)
    np.testing.assert_allclose(O.thomas(ld, d, ud, T), np.linalg.solve(A, T), rtol=1e-12)
    # rows of convective_adjustment! (src/oceananigans_nn.jl:15-22): d_i = 1 + a (kappa_i + kappa_{i+1}), d_N = 1 + a kappa_N
    a = 600.0 / 64.0
    kap = (d[-1] - 1) / a
    assert kap in (0.0, 10.0) and np.all(ld <= 0) and np.all(ud <= 0)


def test_stable_column_only_feels_the_surface_flux(nfc):
    """Zero network, stably stratified column, warming flux (Q < 0 cools... here Q > 0 cools the top cell): no adjustment below."""
    T0 = (20.0 + 0.01 * np.arange(32))[None].astype(np.float64)
    out = O.embedded_solve(T0, [-1e-6], np.zeros(24735), 0, O.chain_spec()[1], [0, 1, 0, 1], 256.0, 600.0, 10.0, 1e-3, 10)
    np.testing.assert_allclose(out[0, -1, :-1], T0[0, :-1], atol=1e-12)            # warming keeps the column stable: interior untouched
    np.testing.assert_allclose(out[0, -1, -1], T0[0, -1] + 10 * 600.0 * 1e-6 / 8.0, rtol=1e-12)


def test_convective_adjustment_mixes_unstable_column(nfc):
    """Cooling from the top: the implicit adjustment keeps the column statically stable (no large negative dT/dz survives)."""
    T0, Q, scal = _columns(nfc, 4, seed=3)
    out = O.embedded_solve(T0, np.abs(Q) + 5e-6, np.zeros(24735), 0, O.chain_spec()[1], [0, 1, 0, 1], 256.0, 600.0, 10.0, 1e-3, 288)
    dTdz = np.diff(out[:, -1], axis=1) / 8.0
    assert dTdz.min() > -2e-4 and (out[:, -1, -1] < T0[:, -1]).all()


@pytest.mark.gpu
@pytest.mark.parametrize("arch,scale", [("dense-default", 0.0), ("dense-default", 1.0), ("conv-2", 1.0), ("dense-deeper", 1.0)])
def test_gpu_embedded_matches_oracle(nfc, gpu_ok, arch, scale):
    B = 37
    T0, Q, scal = _columns(nfc, B, seed=5)
    cw, layers = O.chain_spec(arch)
    theta = O.glorot_theta(cw, layers, 1, scale * 0.3)
    h = nfc.Handle(cw, layers, 1, np.float32([0, 1]))
    h.set_weights(theta)
    K = scal[3] / scal[1] * 691200.0 / 256.0 * 10.0                 # the reference's non-dimensionalisation (src/oceananigans_nn.jl:195)
    out = h.embedded_solve(T0, Q, scal, 256.0, 600.0, K, 1e-3, 144, save_every=12)
    ref = O.embedded_solve(T0, Q, theta, cw, layers, scal, 256.0, 600.0, K, 1e-3, 144, save_every=12)
    assert out.shape == (B, 13, 32)
    np.testing.assert_array_equal(out[:, 0], T0)
    # compare the CHANGE of the profile (temperatures are ~20 C, changes ~0.1 C: a plain relative error would hide everything)
    assert rel_l2(out[:, 1:] - T0[:, None], ref[:, 1:] - T0[:, None]).max() < 2e-3


@pytest.mark.gpu
@pytest.mark.parametrize("tc", ["0", "1"])
def test_gpu_embedded_unstable_bottom_cell(nfc, gpu_ok, monkeypatch, tc):
    """The reference's matrix keeps kappa_1 on the diagonal of its first row without a sub-diagonal (src/oceananigans_nn.jl:19-22): when the
    deepest cell is unstable the operator does not preserve constants, and a solve carried in scaled units must add the affine term.
    Columns with a warm bottom cell against the float64 oracle, which solves in physical units like the reference."""
    monkeypatch.setenv("NFC_TC", tc)
    B = 133
    T0, Q, scal = _columns(nfc, B, seed=11)
    T0 = T0.copy()
    T0[:, 0] += 0.5                                                   # deepest cell warmer than the one above: kappa_1 = K
    cw, layers = O.chain_spec()
    theta = O.glorot_theta(cw, layers, 1, 0.3)
    h = nfc.Handle(cw, layers, 1, np.float32([0, 1]))
    h.set_weights(theta)
    K = scal[3] / scal[1] * 691200.0 / 256.0 * 10.0
    out = h.embedded_solve(T0, Q, scal, 256.0, 600.0, K, -1e-3, 48, save_every=12)
    ref = O.embedded_solve(T0, Q, theta, cw, layers, scal, 256.0, 600.0, K, -1e-3, 48, save_every=12)
    assert rel_l2(out[:, 1:] - T0[:, None], ref[:, 1:] - T0[:, None]).max() < 2e-3
    # without the affine term the deep cells are off by O(mu a kappa_1) ~ kelvins: make sure the test would see it
    assert np.abs(ref[:, 1, 0] - T0[:, 0]).min() > 1e-3


@pytest.mark.gpu
def test_gpu_embedded_api_mirror(nfc, gpu_ok):
    syn = nfc.synthetic
    Ts, wTs = syn.reference_scalings()
    d = syn.make_columns("single")
    zf = -256.0 + np.arange(33) * 8.0
    ds = nfc.ColumnDataset(T=np.repeat(Ts.inv(d["T0"][0])[:, None], 1153, axis=1), wT=np.zeros((33, 1153)), zc=syn.cell_centres(), zf=zf,
                           times=np.arange(1153) * 600.0, metadata={"temperature_flux": 2.5e-5, "dθdz_deep": 1e-3})
    sol = nfc.oceananigans_convective_adjustment(ds, dt=600.0, K=10.0)
    T = sol.T
    assert T.shape == (32, 1153) and np.isfinite(T).all()
    assert T[-1, -1] < T[-1, 0] and np.diff(T[:, -1]).min() > -1e-3      # surface cooled, column adjusted to (near) neutral
    # the reference's second output (diagnose_wT_NN): zero network => wT = [0; 0...; Q] - min(0, K dT/dz) on the interior faces
    assert sol.wT.shape == (33, 1153)
    dTdz = np.diff(T.astype(np.float64), axis=0) / 8.0
    ref = np.zeros((33, 1153)); ref[-1] = 2.5e-5; ref[1:-1] -= np.minimum(0.0, 10.0 * dTdz)
    assert np.abs(sol.wT - ref).max() <= 1e-5 * max(np.abs(ref).max(), 2.5e-5) + 1e-9
    NN = nfc.named_chain("dense-default", seed=0)
    for p in NN.params():
        p *= 1e-5
    sol2 = nfc.oceananigans_convective_adjustment_with_neural_network(ds, NN, Ts, wTs, save_every=9)
    assert sol2.T.shape == (32, 129) and np.isfinite(sol2.T).all() and sol2.wT.shape == (33, 129) and np.isfinite(sol2.wT).all()
    # against the oracle's flux (scaled units) for a few saved profiles
    cw, layers = O.chain_spec()
    theta, _ = nfc.destructure(NN)
    Knd = wTs.sigma / Ts.sigma * 691200.0 / 256.0 * 10.0
    Tsc = Ts(sol2.T[:, ::32].T)
    colp = np.float64([[(0.0 - wTs.mu) / wTs.sigma, (2.5e-5 - wTs.mu) / wTs.sigma, Knd * Ts.sigma / (256.0 * wTs.sigma)]] * Tsc.shape[0])
    wref = wTs.inv(O.total_flux(Tsc, theta, cw, layers, colp, np.float64))
    assert np.abs(sol2.wT[:, ::32].T - wref).max() <= 2e-4 * np.abs(wref).max()


@pytest.mark.gpu
def test_gpu_embedded_tensor_core_matches_simt(nfc, gpu_ok, monkeypatch):
    """Default Chain: the Chain of the embedded path runs on the tensor cores (nfc_embedded_tc.cuh) and the tridiagonal systems are
    solved by a per-column Thomas sweep; the FP32 SIMT kernel uses parallel cyclic reduction.  Two float32 tridiagonal algorithms
    and the FP32-faithful rounding of the network output over 288 steps (with the adjustment's active set flipping at neutral
    stratification): a few 1e-4 of the profile CHANGE; both are within 2e-3 of the float64 oracle (test above).  Ragged tile
    (300 columns = 2 full tiles + 44), several saves."""
    B = 300
    T0, Q, scal = _columns(nfc, B, seed=9)
    cw, layers = O.chain_spec()
    theta = O.glorot_theta(cw, layers, 1, 0.3)
    K = scal[3] / scal[1] * 691200.0 / 256.0 * 10.0
    out = {}
    for tc in ("0", "1"):
        monkeypatch.setenv("NFC_TC", tc)
        h = nfc.Handle(cw, layers, 1, np.float32([0, 1]))
        h.set_weights(theta)
        out[tc] = h.embedded_solve(T0, Q, scal, 256.0, 600.0, K, 1e-3, 288, save_every=24)
    a, b = out["0"], out["1"]
    assert a.shape == b.shape == (B, 13, 32) and np.isfinite(b).all()
    np.testing.assert_array_equal(b[:, 0], T0)
    assert rel_l2(b[:, 1:] - T0[:, None], a[:, 1:] - T0[:, None]).max() < 5e-4
