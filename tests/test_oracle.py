"""CPU tests of the oracle itself: known answers, order conditions, golden vectors, C twin vs NumPy."""
import numpy as np
import pytest

from conftest import golden, rel_l2
from oracle import c_oracle as CO
from oracle import nde_oracle as O

ARCHS = ["dense-default", "dense-wider", "dense-deeper", "conv-2", "conv-4"]


def _data(nfc, kind="training9"):
    return nfc.synthetic.make_columns(kind)


def test_tsit5_order_conditions():
    """The tableau is 5th order, the embedded solution 4th order, the dense output 4th order for every theta."""
    A, b, bt, c, R = O.TSIT5_A, O.TSIT5_B, O.TSIT5_BTILDE, O.TSIT5_C, O.TSIT5_R
    assert np.abs(A.sum(1) - c).max() < 1e-14
    Ac, Ac2, Ac3 = A @ c, A @ c**2, A @ c**3
    AAc = A @ Ac

    def conds(bv, th, upto):
        out = [bv.sum() - th, bv @ c - th**2 / 2, bv @ c**2 - th**3 / 3, bv @ Ac - th**3 / 6, bv @ c**3 - th**4 / 4,
               bv @ (c * Ac) - th**4 / 8, bv @ Ac2 - th**4 / 12, bv @ AAc - th**4 / 24]
        if upto == 5:
            out += [bv @ c**4 - 1 / 5, bv @ (c**2 * Ac) - 1 / 10, bv @ (Ac * Ac) - 1 / 20, bv @ (c * Ac2) - 1 / 15,
                    bv @ Ac3 - 1 / 20, bv @ (c * AAc) - 1 / 30, bv @ (A @ (c * Ac)) - 1 / 40, bv @ (A @ Ac2) - 1 / 60,
                    bv @ (A @ AAc) - 1 / 120]
        return np.abs(out).max()

    assert conds(b, 1.0, 5) < 1e-14
    assert conds(b - bt, 1.0, 4) < 1e-14
    assert abs(bt.sum()) < 1e-15
    for th in np.linspace(0.05, 1.0, 20):
        bth = R @ np.array([th, th**2, th**3, th**4])
        assert conds(bth, th, 4) < 5e-14
    assert np.abs(R.sum(1) - b).max() < 2e-14     # theta = 1 reproduces b


@pytest.mark.parametrize("arch", ARCHS)
def test_param_counts(arch):
    """24735 / 82207 / 41247 / 24610 / 24356."""
    want = {"dense-default": 24735, "dense-wider": 82207, "dense-deeper": 41247, "conv-2": 24610, "conv-4": 24356}
    cw, layers = O.chain_spec(arch)
    assert O.n_params(cw, layers) == want[arch]


def test_destructure_layout_default():
    """SURVEY.md 8 a4 offsets: W1 0..4095, b1 4096.., W2 4224.., b2 20608.., W3 20736.., b3 24704..24734."""
    cw, layers = O.chain_spec()
    theta = np.arange(24735, dtype=np.float64)
    _, dense = O.unpack_theta(theta, cw, layers)
    (W1, b1, _), (W2, b2, _), (W3, b3, _) = dense
    assert W1.shape == (128, 32) and W1[0, 0] == 0 and W1[1, 0] == 1 and W1[0, 1] == 128       # column-major
    assert b1[0] == 4096 and W2[0, 0] == 4224 and b2[0] == 20608 and W3[0, 0] == 20736 and W3[30, 127] == 24703
    assert b3[0] == 24704 and b3[-1] == 24734


@pytest.mark.parametrize("arch", ARCHS)
@pytest.mark.parametrize("K", [0.0, 2.0, 10.0])
def test_heat_budget_telescopes(nfc, arch, K):
    """sum_k dT_k/dt * dz = -pref (top - bottom) for any weights and any K_CA."""
    d = _data(nfc)
    cw, layers = O.chain_spec(arch)
    theta = O.glorot_theta(cw, layers, 3, 1.0)
    colp = d["colp"].astype(np.float64)
    colp[:, 2] = K
    f = O.rhs(d["T0"] + 0.1 * np.random.default_rng(0).standard_normal(d["T0"].shape), theta, cw, layers, colp, d["glob"])
    pref = O.prefactor(d["glob"], np.float64)
    np.testing.assert_allclose(f.sum(1) / 32, -pref * (colp[:, 1] - colp[:, 0]), rtol=1e-10)


def test_zero_network_known_answers(nfc):
    """NN == 0, K_CA = 0: only the two boundary cells change, linearly in time (SURVEY.md 8c v)."""
    d = _data(nfc)
    cw, layers = O.chain_spec()
    theta = np.zeros(O.n_params(cw, layers))
    colp = d["colp"].astype(np.float64)
    colp[:, 2] = 0
    sv = nfc.synthetic.saveat(9).astype(np.float64)
    r = O.tsit5_solve(d["T0"], theta, cw, layers, colp, d["glob"], sv, dtype=np.float64)
    pref = O.prefactor(d["glob"], np.float64)
    T0 = d["T0"].astype(np.float64)
    np.testing.assert_allclose(r["T"][:, :, 1:-1], np.repeat(T0[:, None, 1:-1], 9, 1), atol=1e-12)
    np.testing.assert_allclose(r["T"][:, :, -1], T0[:, None, -1] - pref * colp[:, None, 1] * sv[None] * 32, rtol=1e-9)
    np.testing.assert_allclose(r["T"][:, :, 0], T0[:, None, 0] + pref * colp[:, None, 0] * sv[None] * 32, rtol=1e-9)


def test_stable_profile_has_no_convective_adjustment(nfc):
    """dT/dz >= 0 everywhere => the CA term is exactly zero; and K_CA = 0 equals the free NDE bit for bit."""
    d = _data(nfc)
    cw, layers = O.chain_spec()
    theta = O.glorot_theta(cw, layers, 0, 1.0)
    T = np.sort(d["T0"], axis=1)                                   # monotonically increasing upward
    c0, c2 = d["colp"].copy(), d["colp"].copy()
    c0[:, 2], c2[:, 2] = 0.0, 2.0
    assert np.array_equal(O.rhs(T, theta, cw, layers, c0, d["glob"]), O.rhs(T, theta, cw, layers, c2, d["glob"]))


@pytest.mark.parametrize("arch", ARCHS)
def test_rhs_golden_numpy_and_c(arch):
    g = golden(f"rhs_{arch}.npz")
    cw, layers = O.chain_spec(arch)
    f64 = O.rhs(g["T"], g["theta"], cw, layers, g["colp"], g["glob"], np.float64)
    assert rel_l2(f64, g["dTdt"]).max() < 1e-13
    f32 = O.rhs(g["T"], g["theta"], cw, layers, g["colp"], g["glob"], np.float32)
    assert rel_l2(f32, g["dTdt"]).max() < 1e-5
    fc = CO.rhs(g["T"], g["theta"], cw, layers, g["colp"], g["glob"])
    assert rel_l2(fc, g["dTdt"]).max() < 1e-5          # north_star: per-step RHS within 1e-5 relative in Float32


@pytest.mark.parametrize("tag", ["1e-5", "0p3", "1p0"])
@pytest.mark.parametrize("K", [0, 2])
def test_trajectory_golden(tag, K):
    """float32 oracles (NumPy and C) against the float64 controller-faithful run and the 1e-11 truth: 1e-3 relative."""
    g = golden(f"traj_{tag}_K{K}.npz")
    cw, layers = O.chain_spec()
    theta = O.glorot_theta(cw, layers, int(g["theta_seed"]), float(g["theta_scale"]))
    assert rel_l2(g["T_tsit5_f64"], g["T_truth"]).max() < 1e-3
    c = CO.solve(g["T0"], theta, cw, layers, g["colp"], g["glob"], g["saveat"])
    assert (c["status"] == 0).all()
    assert rel_l2(c["T"], g["T_truth"]).max() < 1e-3
    assert rel_l2(c["T"], g["T_tsit5_f64"]).max() < 1e-3
    # step counts of the float32 twin stay close to the float64 controller's
    assert abs(int(c["naccept"].sum()) - int(g["naccept"].sum())) <= max(3 * c["naccept"].size, 0.05 * g["naccept"].sum())
    assert np.abs(c["naccept"] - g["naccept"]).max() <= max(8, 0.25 * g["naccept"].max())
    if tag == "0p3":
        p = O.tsit5_solve(g["T0"], theta, cw, layers, g["colp"], g["glob"], g["saveat"], dtype=np.float32)
        assert rel_l2(p["T"], g["T_truth"]).max() < 1e-3


@pytest.mark.parametrize("K", [0, 2])
def test_gradient_golden_fixed_step(K):
    """C discrete adjoint vs float64 autograd on a fixed step grid (the well-conditioned setting)."""
    g = golden(f"grad_K{K}.npz")
    cw, layers = O.chain_spec()
    theta = O.glorot_theta(cw, layers, int(g["theta_seed"]), float(g["theta_scale"]))
    r = CO.loss_grad(g["T0"], theta, cw, layers, g["colp"], g["glob"], g["saveat"], g["Ttrue"], fixed_dt=float(g["fixed_dt"]))
    assert abs(r["loss"] - float(g["loss"])) / float(g["loss"]) < 1e-5
    tol = 1e-4 if K == 0 else 1e-3
    assert np.linalg.norm(r["grad"] - g["grad"]) / np.linalg.norm(g["grad"]) < tol
    assert np.linalg.norm(r["dT0"] - g["dT0"]) / np.linalg.norm(g["dT0"]) < tol


def test_gradient_matches_finite_differences(nfc):
    """autograd oracle vs central differences of the float64 loss (smooth setting: K_CA = 0, fixed steps)."""
    d = _data(nfc)
    cw, layers = O.chain_spec()
    theta = O.glorot_theta(cw, layers, 0, 1.0).astype(np.float64)
    sv = nfc.synthetic.saveat(5)
    colp = d["colp"][:2].copy()
    colp[:, 2] = 0
    Ttrue = np.zeros((2, 5, 32))
    L, g, _ = O.loss_and_grad_autograd(d["T0"][:2], theta, cw, layers, colp, d["glob"], sv, Ttrue, fixed_dt=1 / 64)
    w = np.random.default_rng(0).standard_normal(theta.size)

    def loss(th):
        return O.mse(O.tsit5_solve(d["T0"][:2], th, cw, layers, colp, d["glob"], sv, dtype=np.float64, fixed_dt=1 / 64)["T"], Ttrue)

    fd = (loss(theta + 1e-6 * w) - loss(theta - 1e-6 * w)) / 2e-6
    assert abs(fd - g @ w) / abs(fd) < 1e-5


def test_empty_and_ragged_batches(nfc):
    d = _data(nfc)
    cw, layers = O.chain_spec()
    theta = O.glorot_theta(cw, layers, 0, 0.3)
    sv = nfc.synthetic.saveat(3)
    r = CO.solve(d["T0"][:0], theta, cw, layers, d["colp"][:0], d["glob"], sv)
    assert r["T"].shape == (0, 3, 32)
    r1 = CO.solve(d["T0"][:1], theta, cw, layers, d["colp"][:1], d["glob"], sv)
    r9 = CO.solve(d["T0"], theta, cw, layers, d["colp"], d["glob"], sv)
    assert np.array_equal(r1["T"][0], r9["T"][0])                 # columns are independent
    np.testing.assert_array_equal(r9["T"][:, 0], d["T0"])         # first save point is the initial condition


# ---------------------------------------------------------------- RKC2 (stabilised explicit) oracle
def test_rkc_coefficients_are_consistent_and_second_order():
    """Stage 'times' c_j follow the same recurrence as the stages; consistency needs c_s = 1 and second order needs the
    stability polynomial P(z) = 1 + z + z^2/2 + O(z^3): check P on the scalar test equation y' = lambda y."""
    tab = O.rkc_coefficients()
    for m in (2, 3, 5, 9, 20, 64):
        c = tab[m]
        cs = [0.0, c[1, 3]]
        for j in range(2, m + 1):
            assert abs(c[j, 0] + c[j, 1] + c[j, 2] - 1.0) < 1e-14
            cs.append(c[j, 0] * cs[j - 1] + c[j, 1] * cs[j - 2] + c[j, 3] + c[j, 4])
        assert abs(cs[-1] - 1.0) < 1e-12
        for z in (-1e-3, 1e-3):
            y0, y1 = 1.0, 1.0 + z * c[1, 3]
            for j in range(2, m + 1):
                y0, y1 = y1, c[j, 0] * y1 + c[j, 1] * y0 + c[j, 2] + (c[j, 3] * z * y1 + c[j, 4] * z)
            assert abs(y1 - (1 + z + z * z / 2)) < 2e-9          # third-order remainder
        # damped stability along the negative real axis: beta(m) ~ 0.653 (m^2 - 1), which is what the stage-count rule assumes
        for z in np.linspace(-0.653 * (m * m - 1), -0.01, 400):
            y0, y1 = 1.0, 1.0 + z * c[1, 3]
            for j in range(2, m + 1):
                y0, y1 = y1, c[j, 0] * y1 + c[j, 1] * y0 + c[j, 2] + (c[j, 3] * z * y1 + c[j, 4] * z)
            assert abs(y1) <= 1.0 + 1e-9


def test_rkc_oracle_converges_and_beats_tsit5_on_stiff_adjustment(nfc):
    syn = nfc.synthetic
    d = syn.make_columns("training9")
    d["colp"][:, 2] = 10.0
    sv = syn.saveat(17)
    cw, layers = O.chain_spec()
    theta = O.glorot_theta(cw, layers, 0, 0.3)
    truth = O.truth_solve(d["T0"][:3], theta, cw, layers, d["colp"][:3], d["glob"], sv)
    for dt_ in (np.float32, np.float64):
        r = O.rkc_solve(d["T0"][:3], theta, cw, layers, d["colp"][:3], d["glob"], sv, dtype=dt_)
        assert (r["status"] == 0).all()
        assert rel_l2(r["T"], truth).max() < 2e-4
    t = O.tsit5_solve(d["T0"][:3], theta, cw, layers, d["colp"][:3], d["glob"], sv, dtype=np.float32)
    assert r["nfe"].sum() < 0.5 * 6 * (t["naccept"] + t["nreject"]).sum()
    assert O.rkc_max_stages(1e-4, np.float32) == 9 and O.rkc_max_stages(1e-4, np.float64) == O.RKC_MMAX
