"""GPU parity tests of the forward path (through the C ABI): CUDA kernels vs the CPU oracle (restatement of the
reference, parity unpinned -- SURVEY.md 8c) on the same seeded inputs, golden vectors, and size-independent
properties at BASELINE sizes.  Tolerances are the ones BASELINE.json's north_star states."""
import numpy as np
import pytest

from conftest import golden, rel_l2
from oracle import c_oracle as CO
from oracle import nde_oracle as O

pytestmark = pytest.mark.gpu
ARCHS = ["dense-default", "dense-wider", "dense-deeper", "conv-2", "conv-4"]
RHS_TOL = 1e-5      # per-step RHS, Float32, relative (norm-wise per column)
TRAJ_TOL = 1e-3     # full-trajectory temperature, relative


def _handle(nfc, arch="dense-default", nde_type=1, saveat=None, **kw):
    cw, layers = O.chain_spec(arch)
    return nfc.Handle(cw, layers, nde_type, nfc.synthetic.saveat(33) if saveat is None else saveat, **kw), cw, layers


@pytest.mark.parametrize("arch", ARCHS)
def test_rhs_golden(nfc, gpu_ok, arch):
    g = golden(f"rhs_{arch}.npz")
    h, cw, layers = _handle(nfc, arch)
    h.set_weights(g["theta"])
    f = h.rhs(g["T"], g["colp"], g["glob"])
    assert rel_l2(f, g["dTdt"]).max() < RHS_TOL


@pytest.mark.parametrize("arch", ARCHS)
@pytest.mark.parametrize("B", [1, 7, 8, 9, 1000, 4099])
def test_rhs_vs_oracle_ragged(nfc, gpu_ok, arch, B):
    d = nfc.synthetic.make_columns("ocean", B=B, seed=B)
    rng = np.random.default_rng(B)
    T = (d["T0"] + 0.3 * rng.standard_normal(d["T0"].shape)).astype(np.float32)
    colp = d["colp"].copy()
    colp[:, 2] = rng.uniform(0, 10, B)
    h, cw, layers = _handle(nfc, arch)
    theta = O.glorot_theta(cw, layers, 5, 1.0)
    theta[-31:] = 0.1
    h.set_weights(theta)
    f = h.rhs(T, colp, d["glob"])
    ref = O.rhs(T, theta, cw, layers, colp, d["glob"], np.float64)
    assert rel_l2(f, ref).max() < RHS_TOL
    # heat-budget identity holds for the kernel too (no oracle needed)
    pref = O.prefactor(d["glob"], np.float64)
    np.testing.assert_allclose(f.astype(np.float64).sum(1) / 32, -pref * (colp[:, 1].astype(np.float64) - colp[:, 0]), rtol=2e-3, atol=1e-4)


def test_rhs_free_convection_ignores_kca(nfc, gpu_ok):
    """FreeConvectionNDE == ConvectiveAdjustmentNDE with K_CA = 0, bit for bit."""
    d = nfc.synthetic.make_columns("ocean", B=64, seed=3)
    cw, layers = O.chain_spec()
    theta = O.glorot_theta(cw, layers, 1, 1.0)
    hf, _, _ = _handle(nfc, nde_type=0)
    hc, _, _ = _handle(nfc, nde_type=1)
    hf.set_weights(theta)
    hc.set_weights(theta)
    c0 = d["colp"].copy()
    c0[:, 2] = 0
    assert np.array_equal(hf.rhs(d["T0"], d["colp"], d["glob"]), hc.rhs(d["T0"], c0, d["glob"]))


@pytest.mark.parametrize("tag", ["1e-5", "0p3", "1p0"])
@pytest.mark.parametrize("K", [0, 2])
def test_trajectory_golden(nfc, gpu_ok, tag, K):
    g = golden(f"traj_{tag}_K{K}.npz")
    h, cw, layers = _handle(nfc, saveat=g["saveat"])
    theta = O.glorot_theta(cw, layers, int(g["theta_seed"]), float(g["theta_scale"]))
    h.set_weights(theta)
    r = h.solve(g["T0"], g["colp"], g["glob"])
    assert (r["status"] == 0).all()
    assert rel_l2(r["T"], g["T_truth"]).max() < TRAJ_TOL
    assert rel_l2(r["T"], g["T_tsit5_f64"]).max() < TRAJ_TOL
    np.testing.assert_array_equal(r["T"][:, 0], g["T0"])
    assert abs(int(r["naccept"].sum()) - int(g["naccept"].sum())) <= max(27, 0.05 * g["naccept"].sum())


@pytest.mark.parametrize("arch", ARCHS)
def test_solve_vs_c_oracle_all_chains(nfc, gpu_ok, arch):
    """Same float32 algorithm on CPU and GPU: trajectories within tolerance, step counts close."""
    B = 203
    d = nfc.synthetic.make_columns("ocean", B=B, seed=11)
    sv = nfc.synthetic.saveat(17)
    h, cw, layers = _handle(nfc, arch, saveat=sv)
    theta = O.glorot_theta(cw, layers, 2, 0.3)
    h.set_weights(theta)
    r = h.solve(d["T0"], d["colp"], d["glob"])
    c = CO.solve(d["T0"], theta, cw, layers, d["colp"], d["glob"], sv, nthreads=8)
    assert (r["status"] == 0).all() and (c["status"] == 0).all()
    assert rel_l2(r["T"], c["T"]).max() < TRAJ_TOL
    tot_g, tot_c = int((r["naccept"] + r["nreject"]).sum()), int((c["naccept"] + c["nreject"]).sum())
    assert abs(tot_g - tot_c) <= 0.03 * tot_c
    cnt = h.counters()
    assert cnt["steps_accepted"] == int(r["naccept"].sum()) and cnt["steps_rejected"] == int(r["nreject"].sum())


def test_solve_fixed_step_matches_oracle_tightly(nfc, gpu_ok):
    """adaptive=false removes controller divergence: GPU and C oracle run the identical step grid."""
    B = 64
    d = nfc.synthetic.make_columns("ocean", B=B, seed=5)
    sv = nfc.synthetic.saveat(9)
    h, cw, layers = _handle(nfc, saveat=sv, fixed_dt=1 / 256)
    theta = O.glorot_theta(cw, layers, 2, 1.0)
    h.set_weights(theta)
    r = h.solve(d["T0"], d["colp"], d["glob"])
    c = CO.solve(d["T0"], theta, cw, layers, d["colp"], d["glob"], sv, fixed_dt=1 / 256, nthreads=8)
    assert (r["naccept"] == 256).all() and (r["nreject"] == 0).all()
    assert rel_l2(r["T"], c["T"]).max() < 2e-5


def test_solve_columns_are_independent_and_order_free(nfc, gpu_ok):
    """Slot refill / queue order must not leak between columns: permuting the batch permutes the output bit for bit."""
    B = 777
    d = nfc.synthetic.make_columns("ocean", B=B, seed=9)
    sv = nfc.synthetic.saveat(5)
    h, cw, layers = _handle(nfc, saveat=sv)
    h.set_weights(O.glorot_theta(cw, layers, 2, 0.3))
    r = h.solve(d["T0"], d["colp"], d["glob"])
    perm = np.random.default_rng(0).permutation(B)
    rp = h.solve(d["T0"][perm], d["colp"][perm], d["glob"])
    assert np.array_equal(rp["T"], r["T"][perm])
    assert np.array_equal(rp["naccept"], r["naccept"][perm])
    h.set_option("fwd_small_max", 0)                     # same (tensor-core) kernel for a batch of one
    r1 = h.solve(d["T0"][5:6], d["colp"][5:6], d["glob"])
    assert np.array_equal(r1["T"][0], r["T"][5])
    # small batches run one column per CTA (nfc_forward_small.cuh, FP32): its own arithmetic, order-free as well, and within the
    # trajectory tolerance of the tensor-core kernel and of the oracle
    h.set_option("fwd_small_max", 148)
    sub = np.arange(100)
    s1 = h.solve(d["T0"][sub], d["colp"][sub], d["glob"])
    s2 = h.solve(d["T0"][sub[::-1]], d["colp"][sub[::-1]], d["glob"])
    assert np.array_equal(s2["T"], s1["T"][::-1]) and np.array_equal(s2["naccept"], s1["naccept"][::-1])
    assert not np.array_equal(s1["T"], r["T"][sub]) and (s1["status"] == 0).all()
    assert rel_l2(s1["T"], r["T"][sub]).max() < 2 * TRAJ_TOL
    c = CO.solve(d["T0"][:24], O.glorot_theta(cw, layers, 2, 0.3), cw, layers, d["colp"][:24], d["glob"], sv)
    assert rel_l2(s1["T"][:24], c["T"]).max() < TRAJ_TOL


def test_solve_edge_cases(nfc, gpu_ok):
    d = nfc.synthetic.make_columns("training9")
    h, cw, layers = _handle(nfc, saveat=nfc.synthetic.saveat(1153))
    theta = O.glorot_theta(cw, layers, 0, 1e-5)
    h.set_weights(theta)
    r = h.solve(d["T0"][:0], d["colp"][:0], d["glob"])            # empty batch
    assert r["T"].shape == (0, 1153, 32)
    r = h.solve(d["T0"], d["colp"], d["glob"])                    # config 1 shape: 1153 saves, x1e-5 weights
    assert r["T"].shape == (9, 1153, 32) and np.isfinite(r["T"]).all() and (r["status"] == 0).all()
    c = CO.solve(d["T0"], theta, cw, layers, d["colp"], d["glob"], nfc.synthetic.saveat(1153))
    assert rel_l2(r["T"], c["T"]).max() < TRAJ_TOL
    with pytest.raises(nfc.NfcError, match="elements"):           # wrong theta length is a usage error
        h.set_weights(theta[:-1])
    # NaN input: reported per column in status[], never as a call failure (reference: retcode warnings)
    Tbad = d["T0"].copy()
    Tbad[3, 7] = np.nan
    r = h.solve(Tbad, d["colp"], d["glob"])
    assert r["status"][3] == 2 and (np.delete(r["status"], 3) == 0).all()


def test_diagnose_flux_matches_oracle(nfc, gpu_ok):
    d = nfc.synthetic.make_columns("training9")
    sv = nfc.synthetic.saveat(9)
    h, cw, layers = _handle(nfc, saveat=sv)
    theta = O.glorot_theta(cw, layers, 0, 1.0)
    h.set_weights(theta)
    r = h.solve(d["T0"], d["colp"], d["glob"])
    wT = h.diagnose_flux(r["T"], d["colp"])
    ref = np.stack([O.total_flux(r["T"][:, n], theta, cw, layers, d["colp"], np.float64) for n in range(9)], axis=1)
    assert rel_l2(wT, ref).max() < RHS_TOL


def test_reference_api_mirror_single_column(nfc, gpu_ok):
    """solve_nde(nde, NN, T0, alg, nde_params) |> Array is Nz x n_save, and the ds-method un-scales (src/solve.jl)."""
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
    assert nde.saveat.size == 129 and abs(float(nde.tspan[1]) - 1.0) < 1e-7
    T = nfc.solve_nde(nde, NN, Ts(ds.T[:, 0]), nfc.Tsit5(), p)
    assert T.shape == (32, 129)
    theta, _ = nfc.destructure(NN)
    cw, layers = O.chain_spec()
    colp = np.array([[p[0], p[1], p[6]]], dtype=np.float32)
    c = CO.solve(Ts(ds.T[:, 0])[None], theta, cw, layers, colp, p[2:6], nde.saveat)
    assert rel_l2(T.T[None], c["T"]).max() < TRAJ_TOL
    with pytest.raises(ValueError):                                   # 6-vector into the CA type
        nfc.solve_nde(nde, NN, Ts(ds.T[:, 0]), nfc.Tsit5(), p[:6])
    out = nfc.solve_nde(ds, NN, nfc.ConvectiveAdjustmentNDE, p, nfc.Tsit5(), Ts, wTs)
    assert out["T"].shape == (32, 1153) and out["wT"].shape == (33, 1153)
    np.testing.assert_allclose(out["T"][:, 0], ds.T[:, 0], rtol=1e-5)


def test_full_size_properties_65536(nfc, gpu_ok):
    """BASELINE config 2 (65,536 columns, 129 saves): heat budget of every column at every save time
    (integral of T changes by -pref (top-bottom) t), first save == T0, all statuses ok."""
    import torch
    B = 65536
    d = nfc.synthetic.make_columns("ocean", B=B, seed=1)
    sv = nfc.synthetic.saveat(129)
    h, cw, layers = _handle(nfc, saveat=sv)
    h.set_weights(O.glorot_theta(cw, layers, 0, 0.3))
    dev = torch.device("cuda:0")
    T0, colp = (torch.from_numpy(d[k]).to(dev) for k in ("T0", "colp"))
    glob = d["glob"]
    Tout = torch.empty((B, 129, 32), device=dev)
    na, nr, st = (torch.zeros(B, dtype=torch.int32, device=dev) for _ in range(3))
    h.solve_dev(T0, colp, glob, Tout, na, nr, st)
    h.sync()
    assert int(st.abs().sum()) == 0
    assert torch.equal(Tout[:, 0], T0)
    pref = float(O.prefactor(d["glob"], np.float64))
    mean_T = Tout.double().mean(dim=2)                                          # [B, 129]
    want = mean_T[:, :1] - pref * (colp[:, 1:2] - colp[:, 0:1]).double() * torch.from_numpy(sv).to(dev).double()[None]
    err = (mean_T - want).abs().max().item()
    assert err < 2e-4, err
    cnt = h.counters()
    assert cnt["steps_accepted"] == int(na.sum()) and cnt["columns"] == B
    # a random sample of columns against the C oracle
    idx = np.random.default_rng(0).choice(B, 64, replace=False)
    c = CO.solve(d["T0"][idx], O.glorot_theta(cw, layers, 0, 0.3), cw, layers, d["colp"][idx], d["glob"], sv, nthreads=8)
    assert rel_l2(Tout[torch.from_numpy(idx).to(dev)].cpu().numpy(), c["T"]).max() < TRAJ_TOL


def test_kca_sweep_ensemble(nfc, gpu_ok):
    """BASELINE config 5 (i): the K_CA sweep of figureC_optimize_convective_adjustment.jl:50-59 -- 100 log-spaced K_CA in
    [1e-3, 10] x the 9 training columns as ONE batch with per-column K_CA, zero-weight network and Glorot weights."""
    d = nfc.synthetic.make_columns("training9")
    K = np.logspace(-3, 1, 100).astype(np.float32)
    T0 = np.tile(d["T0"], (100, 1))
    colp = np.tile(d["colp"], (100, 1))
    colp[:, 2] = np.repeat(K, 9)
    sv = nfc.synthetic.saveat(9)
    cw, layers = O.chain_spec()
    h, _, _ = _handle(nfc, saveat=sv)
    for theta in (np.zeros(24735, np.float32), O.glorot_theta(cw, layers, 0, 0.3)):
        h.set_weights(theta)
        r = h.solve(T0, colp, d["glob"])
        c = CO.solve(T0, theta, cw, layers, colp, d["glob"], sv, nthreads=8)
        assert (r["status"] == 0).all()
        assert rel_l2(r["T"], c["T"]).max() < TRAJ_TOL
    # larger K_CA = stiffer = more steps (figure C's runtime curve)
    steps = (r["naccept"] + r["nreject"]).reshape(100, 9).sum(1)
    assert steps[-1] > 3 * steps[0]


@pytest.mark.parametrize("split", ["fp16", "bf16"])
def test_tensor_core_splits_are_fp32_faithful(nfc, gpu_ok, monkeypatch, split):
    """Both tensor-core split schemes (FP16x2 with power-of-two operand scaling = default, BF16x3) meet the 1e-5 RHS bar with
    a large margin, for O(1) weights, the reference's x1e-5 weights and x100 weights (scaling / exponent-range check)."""
    monkeypatch.setenv("NFC_TC", "1")
    monkeypatch.setenv("NFC_TC_SPLIT", split)
    g = golden("rhs_dense-default.npz")
    cw, layers = O.chain_spec()
    for scale in (1.0, 1e-5, 100.0):
        theta = (g["theta"] * scale).astype(np.float32)
        h = nfc.Handle(cw, layers, 1, nfc.synthetic.saveat(9))
        h.set_weights(theta)
        f = h.rhs(g["T"], g["colp"], g["glob"])
        ref = O.rhs(g["T"], theta, cw, layers, g["colp"], g["glob"], np.float64)
        assert rel_l2(f, ref).max() < 2e-6


def test_tensor_core_out_of_range_input_is_flagged_or_resolved(nfc, gpu_ok, monkeypatch):
    """FP16x2 assumes |T_scaled| <= 128 (operand scaling).  The kernel never guesses: a column beyond that ends as status 5 (ST_RANGE), its
    neighbours untouched (`range_fallback` 0).  By default the host-buffer call re-solves such a column on the BF16x3 split (fp32's
    exponent range), so the caller sees what the FP32 kernel integrates, as the reference would."""
    monkeypatch.setenv("NFC_TC", "1")
    monkeypatch.setenv("NFC_TC_SPLIT", "fp16")
    d = nfc.synthetic.make_columns("ocean", B=200, seed=8)
    sv = nfc.synthetic.saveat(5)
    h, cw, layers = _handle(nfc, saveat=sv)
    h.set_weights(O.glorot_theta(cw, layers, 0, 0.3))
    h.set_option("fwd_small_max", 0)           # a batch this small would run the FP32 one-column-per-CTA kernel, which has no range bound
    ok = h.solve(d["T0"], d["colp"], d["glob"])
    T0 = d["T0"].copy()
    T0[17] *= 120.0                            # |T_scaled| up to ~ 167: outside the FP16x2 operand range, harmless in fp32
    assert np.abs(T0[17]).max() > 128
    keep = np.arange(200) != 17
    h.set_option("range_fallback", 0)
    r = h.solve(T0, d["colp"], d["glob"])
    assert r["status"][17] == 5
    assert (r["status"][keep] == 0).all() and np.array_equal(r["T"][keep], ok["T"][keep])
    h.set_option("range_fallback", 1)
    r = h.solve(T0, d["colp"], d["glob"])
    assert (r["status"][keep] == 0).all() and np.array_equal(r["T"][keep], ok["T"][keep])
    h.set_option("tc", 0)                      # the FP32 SIMT kernel on the same inputs
    s = h.solve(T0, d["colp"], d["glob"])
    assert r["status"][17] == s["status"][17] == 0
    assert rel_l2(r["T"][17:18], s["T"][17:18]).max() < TRAJ_TOL
    assert abs(int(r["naccept"][17]) - int(s["naccept"][17])) <= 0.05 * s["naccept"][17] + 2
    # loss + gradient: the whole call is repeated on the FP32 kernels when a column left the range (a batch this small would run the FP32
    # one-column-per-CTA forward kernel anyway, which has no range bound: switched off to reach the tensor-core RECORD pass)
    h.set_option("tc", 1)
    h.set_option("fwd_small_max", 0)
    Ttrue = np.ascontiguousarray(ok["T"] + 0.05, dtype=np.float32)
    g = h.loss_grad(T0[:32], d["colp"][:32], d["glob"], Ttrue[:32])
    h.set_option("tc", 0)
    g0 = h.loss_grad(T0[:32], d["colp"][:32], d["glob"], Ttrue[:32])
    assert (g["status"] == 0).all() and g["loss"] == g0["loss"] and np.array_equal(g["grad"], g0["grad"])
    h.set_option("tc", 1)
    h.set_option("range_fallback", 0)
    g1 = h.loss_grad(T0[:32], d["colp"][:32], d["glob"], Ttrue[:32])
    assert g1["status"][17] == 5 and (np.delete(g1["status"], 17) == 0).all()


@pytest.mark.parametrize("B", [1, 127, 300, 5000])
def test_two_tile_kernel_matches_one_tile_kernel_bit_for_bit(nfc, gpu_ok, monkeypatch, B):
    """nfc_tc2.cuh (two 128-column tiles in flight per CTA, the large-batch kernel) performs the arithmetic of the one-tile
    tensor-core kernel operation for operation: trajectories, step counts and status agree bit for bit, ragged sizes included
    (one tile empty, both partly filled, several refills per slot), with an out-of-range column flagged as in the one-tile kernel."""
    monkeypatch.setenv("NFC_TC", "1")
    monkeypatch.setenv("NFC_TC_SPLIT", "fp16")
    d = nfc.synthetic.make_columns("ocean", B=B, seed=21)
    T0 = d["T0"].copy()
    if B > 200:
        T0[131] *= 1.0e4                      # |T_scaled| > 128: must end as status 5 (ST_RANGE) in both kernels
    sv = nfc.synthetic.saveat(33)
    res = []
    for tiles in ("1", "2"):
        monkeypatch.setenv("NFC_TC_TILES", tiles)
        h, cw, layers = _handle(nfc, saveat=sv)
        h.set_option("range_fallback", 0)     # the kernels' own answer, not the host call's BF16x3 re-solve
        h.set_weights(O.glorot_theta(cw, layers, 0, 0.3))
        res.append(h.solve(T0, d["colp"], d["glob"]))
    a, b = res
    assert np.array_equal(a["status"], b["status"]) and np.array_equal(a["naccept"], b["naccept"]) and np.array_equal(a["nreject"], b["nreject"])
    if B > 200:
        assert b["status"][131] == 5
    good = b["status"] == 0
    assert good.sum() >= B - 1
    assert np.array_equal(a["T"][good].view(np.uint32), b["T"][good].view(np.uint32))
    # and against the oracle at the trajectory tolerance
    k = min(B, 16)
    theta = O.glorot_theta(cw, layers, 0, 0.3)
    ref = CO.solve(d["T0"][:k], theta, cw, layers, d["colp"][:k], d["glob"], sv)
    assert rel_l2(b["T"][:k], ref["T"]).max() < TRAJ_TOL


def test_two_tile_kernel_edge_cases(nfc, gpu_ok, monkeypatch):
    """The two-tile kernel at the seams of its tiling: exactly one / two full tiles and one column more, fixed steps, the free NDE
    (K_CA ignored), a NaN column, a tiny max_iters (status 4 for everybody) -- always the one-tile kernel's answer bit for bit."""
    monkeypatch.setenv("NFC_TC", "1")
    monkeypatch.setenv("NFC_TC_SPLIT", "fp16")
    sv = nfc.synthetic.saveat(17)
    cw, layers = O.chain_spec()
    theta = O.glorot_theta(cw, layers, 3, 0.3)

    def both(B, nde_type=1, bad=None, **kw):
        d = nfc.synthetic.make_columns("ocean", B=B, seed=40 + B % 7)
        T0 = d["T0"].copy()
        if bad is not None:
            T0[bad, 5] = np.nan
        out = []
        for tiles in ("1", "2"):
            monkeypatch.setenv("NFC_TC_TILES", tiles)
            h = nfc.Handle(cw, layers, nde_type, sv, **kw)
            h.set_weights(theta)
            out.append(h.solve(T0, d["colp"], d["glob"]))
        a, b = out
        for k in ("status", "naccept", "nreject"):
            assert np.array_equal(a[k], b[k]), (B, k)
        ok = b["status"] == 0
        assert np.array_equal(a["T"][ok].view(np.uint32), b["T"][ok].view(np.uint32)), B
        return b

    for B in (128, 129, 256, 257):
        r = both(B)
        assert (r["status"] == 0).all()
    r = both(200, fixed_dt=1.0 / 64)
    ok = r["status"] == 0                      # a fixed step of 1/64 is beyond the stability limit of some columns: those end as NaN
    assert ok.sum() > 100 and (r["naccept"][ok] == 64).all() and (r["nreject"][ok] == 0).all()
    both(150, nde_type=0)
    r = both(300, bad=140)
    assert r["status"][140] == 2 and (np.delete(r["status"], 140) == 0).all()
    r = both(140, max_iters=5)
    assert (r["status"] == 1).sum() > 100      # ST_MAXITERS


def test_set_option_selects_kernels(nfc, gpu_ok):
    """nfc_set_option (include/nfc.h): the kernel-selection knobs have an API surface, not only environment variables."""
    syn = nfc.synthetic
    d = syn.make_columns("ocean", B=300, seed=9)
    cw, layers = O.chain_spec("dense-default")
    theta = O.glorot_theta(cw, layers, 0, 0.3)
    h = nfc.Handle(cw, layers, 1, syn.saveat(17))
    h.set_weights(theta)
    h.set_option("tc_tiles", 1)                # the one-tile tensor-core kernel (300 columns would run one column per CTA: fwd_small_max)
    a = h.solve(d["T0"], d["colp"], d["glob"])
    h.set_option("tc", 0)                      # FP32 SIMT kernels
    b = h.solve(d["T0"], d["colp"], d["glob"])
    h.set_option("tc", 1)
    h.set_option("tc_tiles", 2)                # two tiles in flight: bit-identical to one tile
    c = h.solve(d["T0"], d["colp"], d["glob"])
    assert (a["status"] == 0).all() and (b["status"] == 0).all()
    assert rel_l2(a["T"], b["T"]).max() < TRAJ_TOL and not np.array_equal(a["T"], b["T"])
    assert np.array_equal(a["T"], c["T"])
    with pytest.raises(nfc.NfcError):
        h.set_option("no_such_option", 1)
    with pytest.raises(nfc.NfcError):
        h.set_option("tc_tiles", 3)


@pytest.mark.parametrize("arch", ["conv-2", "conv-4", "dense-deeper"])
def test_other_chains_run_on_the_tensor_cores(nfc, gpu_ok, arch):
    """conv-2 / conv-4 / deeper Chains: nfc_rhs and the Tsit5 forward solve run on the tensor-core kernels (templated weight-image layout,
    Conv front-end evaluated by the column threads, third hidden layer as one more MMA layer); `tc` 0 selects the FP32 SIMT kernels.
    Both against the float64 / C oracle at the north-star tolerances, ragged batch (2 full tiles + 44 columns)."""
    syn = nfc.synthetic
    d = syn.make_columns("ocean", B=300, seed=12)
    cw, layers = O.chain_spec(arch)
    sv = syn.saveat(9)
    for scale in (0.3, 1e-5):
        theta = O.glorot_theta(cw, layers, 5, scale)
        h = nfc.Handle(cw, layers, 1, sv)
        h.set_weights(theta)
        rng = np.random.default_rng(3)
        T = (d["T0"] + 0.3 * rng.standard_normal(d["T0"].shape)).astype(np.float32)
        ref = O.rhs(T, theta, cw, layers, d["colp"], d["glob"], np.float64)
        f_tc = h.rhs(T, d["colp"], d["glob"])
        r_tc = h.solve(d["T0"], d["colp"], d["glob"])
        assert h.counters()["kernel_launches"] >= 1
        h.set_option("tc", 0)
        f_simt = h.rhs(T, d["colp"], d["glob"])
        r_simt = h.solve(d["T0"], d["colp"], d["glob"])
        assert rel_l2(f_tc, ref).max() < 1e-5 and rel_l2(f_simt, ref).max() < 1e-5
        assert not np.array_equal(f_tc, f_simt)                                # two different kernels
        assert (r_tc["status"] == 0).all() and (r_simt["status"] == 0).all()
        # two float32 adaptive solves with different rounding (and therefore step sequences): each within the tolerance of the oracle,
        # up to twice that between them
        assert rel_l2(r_tc["T"], r_simt["T"]).max() < 2 * TRAJ_TOL
        c = CO.solve(d["T0"][:48], theta, cw, layers, d["colp"][:48], d["glob"], sv)
        assert rel_l2(r_tc["T"][:48], c["T"]).max() < TRAJ_TOL and rel_l2(r_simt["T"][:48], c["T"]).max() < TRAJ_TOL
