"""CPU ORACLE -- TEST INFRASTRUCTURE ONLY (never imported by the product path).

A NumPy restatement of the NeuralFreeConvection.jl NDE hot path.  Only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` leg may
import this file; the shipped library (neuralfreeconvection.jl_b200/csrc) never does.

PARITY UNPINNED: the reference (Julia) cannot run in this container and ships no
tests / golden vectors.  Every function below restates a
reference source line (cited) or a published algorithm of an un-vendored Julia
dependency (named with its Manifest.toml pin).  What *is* pinned mathematically:
the Tsit5 tableau and its dense-output polynomials satisfy every Runge-Kutta order
condition (tests/test_oracle.py), and closed-form known answers of the RHS.

Layouts (C-ABI wire formats, see include/nfc.h):
  T      [B][Nz]            (== Julia Nz x B column-major), index 0 = sea floor cell
  colp   [B][3]             (bottom_flux, top_flux, K_CA) already scaled by wT_scaling
  glob   [4]                (sigma_T, sigma_wT, H, tau)
  theta  flat Flux.destructure order (per layer: weight out x in column-major, bias)
  Tout   [B][n_save][Nz]    (== Julia Nz x n_save x B column-major)
"""
from __future__ import annotations

import numpy as np

# ----------------------------------------------------------------------------------
# Tsit5 (Tsitouras 2011) as used by OrdinaryDiffEq 5.60.1 (Manifest.toml:1456) [UPSTREAM]
# ----------------------------------------------------------------------------------
TSIT5_C = np.array([0.0, 0.161, 0.327, 0.9, 0.9800255409045097, 1.0, 1.0])
TSIT5_A = np.zeros((7, 7))
TSIT5_A[1, :1] = [0.161]
TSIT5_A[2, :2] = [-0.008480655492356989, 0.335480655492357]
TSIT5_A[3, :3] = [2.8971530571054935, -6.359448489975075, 4.3622954328695815]
TSIT5_A[4, :4] = [5.325864828439257, -11.748883564062828, 7.4955393428898365, -0.09249506636175525]
TSIT5_A[5, :5] = [5.86145544294642, -12.92096931784711, 8.159367898576159, -0.071584973281401,
                  -0.028269050394068383]
TSIT5_B = np.array([0.09646076681806523, 0.01, 0.4798896504144996, 1.379008574103742,
                    -3.290069515436081, 2.324710524099774, 0.0])
TSIT5_A[6, :] = TSIT5_B
TSIT5_BTILDE = np.array([-0.00178001105222577714, -0.0008164344596567469, 0.007880878010261995,
                         -0.1447110071732629, 0.5823571654525552, -0.45808210592918697, 1.0 / 66.0])
# dense output b_i(theta) = sum_p R[i, p] * theta**(p+1)   (OrdinaryDiffEq tsit5 interpolants)
TSIT5_R = np.array([
    [1.0, -2.763706197274826, 2.9132554618219126, -1.0530884977290216],
    [0.0, 0.13169999999999998, -0.2234, 0.1017],
    [0.0, 3.9302962368947516, -5.941033872131505, 2.490627285651253],
    [0.0, -12.411077166933676, 30.33818863028232, -16.548102889244902],
    [0.0, 37.50931341651104, -88.1789048947664, 47.37952196281928],
    [0.0, -27.896526289197286, 65.09189467479366, -34.87065786149661],
    [0.0, 1.5, -4.0, 2.5]])

# PI step-size controller defaults for Tsit5 in OrdinaryDiffEq 5.x [UPSTREAM]
BETA1, BETA2, GAMMA, QMIN, QMAX, QOLDINIT = 7.0 / 50.0, 2.0 / 25.0, 0.9, 0.2, 10.0, 1e-4
RELTOL_DEFAULT = 1e-4   # src/solve.jl:4
ABSTOL_DEFAULT = 1e-6   # DiffEqBase default [UPSTREAM]
MAXITERS = 100000       # DiffEqBase default [UPSTREAM]

ACT_IDENTITY, ACT_RELU = 0, 1


# ----------------------------------------------------------------------------------
# Chain architectures: train_free_convection_nde.jl:125-156
# ----------------------------------------------------------------------------------
def chain_spec(arch: str = "dense-default", Nz: int = 32):
    """Returns (conv_width, [(n_in, n_out, act), ...]) for the reference's named Chains."""
    if arch == "dense-default":
        return 0, [(Nz, 4 * Nz, ACT_RELU), (4 * Nz, 4 * Nz, ACT_RELU), (4 * Nz, Nz - 1, ACT_IDENTITY)]
    if arch == "dense-wider":
        return 0, [(Nz, 8 * Nz, ACT_RELU), (8 * Nz, 8 * Nz, ACT_RELU), (8 * Nz, Nz - 1, ACT_IDENTITY)]
    if arch == "dense-deeper":
        return 0, [(Nz, 4 * Nz, ACT_RELU), (4 * Nz, 4 * Nz, ACT_RELU), (4 * Nz, 4 * Nz, ACT_RELU),
                   (4 * Nz, Nz - 1, ACT_IDENTITY)]
    if arch in ("conv-2", "conv-4"):
        k = int(arch[-1])
        n0 = Nz - k + 1
        return k, [(n0, 4 * Nz, ACT_RELU), (4 * Nz, 4 * Nz, ACT_RELU), (4 * Nz, Nz - 1, ACT_IDENTITY)]
    raise ValueError(f"Invalid neural network architecture: {arch}")


def n_params(conv_width, layers):
    n = (conv_width + 1) if conv_width else 0
    return n + sum(i * o + o for i, o, _ in layers)


def unpack_theta(theta, conv_width, layers):
    """Flux.destructure layout (Flux 0.12.6 [UPSTREAM]; field names weight/bias corroborated by
    movie3_neural_network_weights.jl:46-47): layers in order, weight (out x in, column-major) then bias."""
    theta = np.asarray(theta)
    off = 0
    conv = None
    if conv_width:
        conv = (theta[off:off + conv_width], theta[off + conv_width])
        off += conv_width + 1
    dense = []
    for n_in, n_out, act in layers:
        W = theta[off:off + n_in * n_out].reshape(n_in, n_out).T  # column-major out x in
        off += n_in * n_out
        b = theta[off:off + n_out]
        off += n_out
        dense.append((W, b, act))
    if off != theta.size:
        raise ValueError(f"theta has {theta.size} elements, Chain needs {off}")
    return conv, dense


def glorot_theta(conv_width, layers, seed=0, scale=1.0, dtype=np.float32):
    """Flux default init: glorot_uniform weights U(+-sqrt(6/(fan_in+fan_out))), zero biases [UPSTREAM];
    `scale` mirrors the x1e-5 at train_free_convection_nde.jl:158-160 (applied to weights and biases)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    parts = []
    if conv_width:
        lim = np.sqrt(6.0 / (conv_width + conv_width))
        parts += [rng.uniform(-lim, lim, conv_width), np.zeros(1)]
    for n_in, n_out, _ in layers:
        lim = np.sqrt(6.0 / (n_in + n_out))
        parts += [rng.uniform(-lim, lim, n_in * n_out), np.zeros(n_out)]
    return (np.concatenate(parts) * scale).astype(dtype)


def nn_forward(theta, conv_width, layers, X):
    """NN(T) for a batch X[B, Nz] -> [B, Nz-1].  Dense: act(W x + b) (Flux [UPSTREAM]);
    Conv((k,1),1=>1,relu) is a true convolution (flipped kernel) (train_free_convection_nde.jl:138-153)."""
    conv, dense = unpack_theta(theta, conv_width, layers)
    H = X
    if conv is not None:
        w, b = conv
        k = w.size
        n0 = H.shape[1] - k + 1
        acc = np.zeros((H.shape[0], n0), dtype=H.dtype)
        for j in range(k):
            acc = acc + w[j].astype(H.dtype) * H[:, (k - 1 - j):(k - 1 - j) + n0]
        H = np.maximum(acc + b.astype(H.dtype), 0)
    for W, b, act in dense:
        H = H @ W.T.astype(H.dtype) + b.astype(H.dtype)
        if act == ACT_RELU:
            H = np.maximum(H, 0)
    return H


# ----------------------------------------------------------------------------------
# RHS: src/free_convection_nde.jl:29-38 and src/convective_adjustment_nde.jl:33-48
# ----------------------------------------------------------------------------------
def prefactor(glob, dtype):
    sT, swT, H, tau = [dtype(x) for x in glob]
    return swT / sT * tau / H          # same operation order as convective_adjustment_nde.jl:47


def total_flux(T, theta, conv_width, layers, colp, dtype=np.float64):
    """F[B, Nz+1] = [bottom; NN(T); top] - min(0, K_CA * dT/dz) (src/solve.jl:35-48).
    Dz^f has zero first/last rows (convective_adjustment.jl:21-29, names swapped)."""
    T = np.asarray(T, dtype=dtype)
    B, Nz = T.shape
    colp = np.asarray(colp, dtype=dtype)
    inv_dz = dtype(Nz)                                   # 1/dz_hat, dz_hat = dz/H = 1/Nz
    F = np.empty((B, Nz + 1), dtype=dtype)
    F[:, 0] = colp[:, 0]
    F[:, Nz] = colp[:, 1]
    F[:, 1:Nz] = nn_forward(np.asarray(theta, dtype=dtype), conv_width, layers, T)
    dTdz = (T[:, 1:] - T[:, :-1]) * inv_dz
    F[:, 1:Nz] -= np.minimum(dtype(0), colp[:, 2:3] * dTdz)
    return F


def rhs(T, theta, conv_width, layers, colp, glob, dtype=np.float64):
    """dT/dt_hat = pref * (-(F[k+1]-F[k])/dz_hat); K_CA = 0 reproduces FreeConvectionNDE."""
    F = total_flux(T, theta, conv_width, layers, colp, dtype)
    Nz = F.shape[1] - 1
    return -prefactor(glob, dtype) * (F[:, 1:] - F[:, :-1]) * dtype(Nz)


# ----------------------------------------------------------------------------------
# Adaptive Tsit5, OrdinaryDiffEq semantics [UPSTREAM]; solve settings src/solve.jl:4
# ----------------------------------------------------------------------------------
def _rms(x):
    return np.sqrt(np.mean(x * x, axis=1))


def initial_dt(f, u0, f0, reltol, abstol, dtmax, dtype):
    """Hairer-Wanner initial step as in OrdinaryDiffEq `ode_determine_initdt` [UPSTREAM]: the exponent divides by the algorithm order (get_current_alg_order = 5 for Tsit5), not by order + 1 as in Hairer's book."""
    sk = abstol + np.abs(u0) * reltol
    d0 = _rms(u0 / sk)
    d1 = _rms(f0 / sk)
    small = (d0 < 1e-5) | (d1 < 1e-5)
    dt0 = np.where(small, dtype(1e-6), dtype(0.01) * d0 / np.where(small, 1, d1)).astype(dtype)
    dt0 = np.minimum(dt0, dtmax)
    f1 = f(u0 + dt0[:, None] * f0)
    d2 = _rms((f1 - f0) / sk) / dt0
    m = np.maximum(d1, d2)
    tiny = m <= 1e-15
    dt1 = np.where(tiny, np.maximum(dtype(1e-6), dt0 * dtype(1e-3)),
                   np.power(dtype(10.0), -(2 + np.log10(np.where(tiny, 1, m))) / 5)).astype(dtype)
    return np.minimum(np.minimum(100 * dt0, dt1), dtmax).astype(dtype)


def tsit5_solve(T0, theta, conv_width, layers, colp, glob, saveat, reltol=RELTOL_DEFAULT,
                abstol=ABSTOL_DEFAULT, dtype=np.float64, fixed_dt=0.0, record_tape=False,
                maxiters=MAXITERS):
    """solve_nde(nde, NN, T0, Tsit5(), nde_params) |> Array (src/solve.jl:1-6), batched over columns,
    every column with its own controller state.  Returns dict(T=[B,n_save,Nz], naccept, nreject, status,
    tape=per-column list of (t, h, u) if record_tape).  `fixed_dt > 0` = adaptive=false, dt=fixed_dt."""
    dt_ = dtype
    u = np.array(T0, dtype=dt_)
    B, Nz = u.shape
    theta = np.asarray(theta, dtype=dt_)
    colp = np.asarray(colp, dtype=dt_)
    saveat = np.asarray(saveat, dtype=dt_)
    n_save = saveat.size
    tf = dt_(saveat[-1])
    reltol, abstol = dt_(reltol), dt_(abstol)
    A, Bv, BT, R = (x.astype(dt_) for x in (TSIT5_A, TSIT5_B, TSIT5_BTILDE, TSIT5_R))

    def f_all(U, idx=None):
        cp = colp if idx is None else colp[idx]
        return rhs(U, theta, conv_width, layers, cp, glob, dt_)

    out = np.zeros((B, n_save, Nz), dtype=dt_)
    t = np.zeros(B, dtype=dt_)
    k1 = f_all(u)
    if fixed_dt > 0:
        dt = np.full(B, fixed_dt, dtype=dt_)
    else:
        dt = initial_dt(f_all, u, k1, reltol, abstol, tf, dt_)
    qold = np.full(B, QOLDINIT, dtype=dt_)
    si = np.zeros(B, dtype=np.int64)
    while True:                                       # save points at t0
        m = (si < n_save)
        m[m] = saveat[si[m]] <= t[m]
        if not m.any():
            break
        out[m, si[m]] = u[m]
        si[m] += 1
    naccept = np.zeros(B, dtype=np.int64)
    nreject = np.zeros(B, dtype=np.int64)
    status = np.zeros(B, dtype=np.int32)
    tape = [[] for _ in range(B)] if record_tape else None
    active = t < tf
    for _ in range(maxiters):
        idx = np.nonzero(active)[0]
        if idx.size == 0:
            break
        ua, ta, ka1 = u[idx], t[idx], k1[idx]
        h = np.minimum(dt[idx], tf - ta)
        last = (ta + h) >= tf * dt_(1 - 1e-6)
        h = np.where(last, tf - ta, h).astype(dt_)
        ks = [ka1]
        for s in range(1, 7):
            acc = np.zeros_like(ua)
            for j in range(s):
                acc = acc + A[s, j] * ks[j]
            y = ua + h[:, None] * acc
            if s == 6:
                unew = y
            ks.append(f_all(y, idx))
        err = np.zeros_like(ua)
        for j in range(7):
            err = err + BT[j] * ks[j]
        err = h[:, None] * err
        EEst = _rms(err / (abstol + np.maximum(np.abs(ua), np.abs(unew)) * reltol))
        bad = ~np.isfinite(EEst)
        if fixed_dt > 0:
            acc_m = np.ones(idx.size, dtype=bool)
            dtnew = h.copy()
            q11 = np.ones_like(h)
        else:
            q11 = np.power(EEst, dt_(BETA1))
            q = q11 / np.power(qold[idx], dt_(BETA2))
            q = np.maximum(dt_(1 / QMAX), np.minimum(dt_(1 / QMIN), q / dt_(GAMMA)))
            acc_m = EEst <= 1
            dtnew = np.where(acc_m, h / q, h / np.minimum(dt_(1 / QMIN), q11 / dt_(GAMMA))).astype(dt_)
        acc_m = acc_m & ~bad
        # accepted columns: dense output for every save point in (t, t+h]
        ia = idx[acc_m]
        if ia.size:
            tn = np.where(last[acc_m], tf, ta[acc_m] + h[acc_m]).astype(dt_)
            ha = h[acc_m]
            ksa = [k[acc_m] for k in ks]
            if record_tape:
                for n, col in enumerate(ia):
                    tape[col].append((float(ta[acc_m][n]), float(ha[n]), ua[acc_m][n].copy()))
            while True:
                m = si[ia] < n_save
                m[m] = saveat[si[ia][m]] <= tn[m]
                if not m.any():
                    break
                cols = ia[m]
                th = ((saveat[si[cols]] - ta[acc_m][m]) / ha[m]).astype(dt_)
                th = np.minimum(th, dt_(1))
                pw = np.stack([th, th * th, th * th * th, th * th * th * th], axis=1)
                bth = pw @ R.T                                     # [n, 7]
                val = ua[acc_m][m].copy()
                inc = np.zeros_like(val)
                for j in range(7):
                    inc = inc + bth[:, j:j + 1] * ksa[j][m]
                out[cols, si[cols]] = val + ha[m][:, None] * inc
                si[cols] += 1
            u[ia] = unew[acc_m]
            t[ia] = tn
            k1[ia] = ks[6][acc_m]
            if fixed_dt <= 0:
                qold[ia] = np.maximum(EEst[acc_m], dt_(QOLDINIT))
            naccept[ia] += 1
        nreject[idx[~acc_m]] += 1
        dt[idx] = dtnew
        status[idx[bad]] = 2                                        # NaN / Inf
        active[idx[bad]] = False
        active[ia] = t[ia] < tf
        tiny = active.copy()
        tiny[tiny] = dt[tiny] < dt_(1e-12)
        status[tiny] = 3                                            # dt underflow
        active[tiny] = False
    status[active] = 1                                              # maxiters
    return dict(T=out, naccept=naccept, nreject=nreject, status=status, tape=tape)


# ----------------------------------------------------------------------------------
# a stabilised explicit integrator in the role ROCK4 plays in production
# (train_free_convection_nde.sh:3, train_free_convection_nde.jl:328, figureC:31).  ROCK4's coefficient
# tables live in OrdinaryDiffEq [UPSTREAM, not vendored] and cannot be restated offline, so this is the
# second-order Runge-Kutta-Chebyshev method of Sommeijer, Shampine & Verwer (RKC, J. Comput. Appl. Math.
# 88, 1998), whose coefficients are closed-form Chebyshev recurrences: same family (explicit, stage
# count chosen per step from a spectral-radius estimate, real-axis stability ~0.65 s^2), lower order.
# Driver logic follows the published RKC algorithm: nonlinear power iteration for the spectral radius
# (re-estimated every 25 steps or after a rejection; made robust to the RHS's kinks, see rkc_solve), first step from a difference quotient, error
# estimate 0.8 (y_n - y_{n+1}) + 0.4 h (f_n + f_{n+1}), predictive step-size controller, cubic Hermite
# dense output at the save times.
# ----------------------------------------------------------------------------------
RKC_MMAX = 64          # table size; the per-solve cap is rkc_max_stages()


def rkc_max_stages(reltol, dtype):
    """Round-off grows like m^2 u inside an RKC step: keep m^2 <= reltol / (10 u) (RKC's own rule)."""
    return int(max(2, min(RKC_MMAX, int(np.sqrt(float(reltol) / (10.0 * float(np.finfo(dtype).eps)))))))


def rkc_coefficients(mmax=RKC_MMAX):
    """tab[m, j] = (mu_j, nu_j, 1 - mu_j - nu_j, mu~_j, gamma~_j) for the m-stage formula, j = 2..m; tab[m, 1, 3] = mu~_1.
        Y_1 = y_n + h mu~_1 f_n,   Y_j = mu_j Y_{j-1} + nu_j Y_{j-2} + (1 - mu_j - nu_j) y_n + h (mu~_j f(Y_{j-1}) + gamma~_j f_n)
    with w0 = 1 + (2/13)/m^2, w1 = T'_m(w0)/T''_m(w0), b_j = T''_j(w0)/T'_j(w0)^2 (b_0 = b_1 = b_2), mu_j = 2 b_j w0/b_{j-1},
    nu_j = -b_j/b_{j-2}, mu~_j = 2 b_j w1/b_{j-1}, gamma~_j = -(1 - b_{j-1} T_{j-1}(w0)) mu~_j.  float64."""
    tab = np.zeros((mmax + 1, mmax + 1, 5))
    for m in range(2, mmax + 1):
        w0 = 1.0 + (2.0 / 13.0) / (m * m)
        T, dT, d2T = [1.0, w0], [0.0, 1.0], [0.0, 0.0]
        for j in range(2, m + 1):
            T.append(2 * w0 * T[j - 1] - T[j - 2])
            dT.append(2 * w0 * dT[j - 1] - dT[j - 2] + 2 * T[j - 1])
            d2T.append(2 * w0 * d2T[j - 1] - d2T[j - 2] + 4 * dT[j - 1])
        w1 = dT[m] / d2T[m]
        b = [0.0, 0.0] + [d2T[j] / dT[j] ** 2 for j in range(2, m + 1)]
        b[0] = b[1] = b[2]
        tab[m, 1, 3] = b[1] * w1
        for j in range(2, m + 1):
            mu, nu, mus = 2 * b[j] * w0 / b[j - 1], -b[j] / b[j - 2], 2 * b[j] * w1 / b[j - 1]
            tab[m, j] = (mu, nu, 1.0 - mu - nu, mus, -(1.0 - b[j - 1] * T[j - 1]) * mus)
    return tab


def rkc_solve(T0, theta, conv_width, layers, colp, glob, saveat, reltol=RELTOL_DEFAULT, abstol=ABSTOL_DEFAULT,
              dtype=np.float32, maxiters=MAXITERS):
    """Adaptive RKC2 for every column (own controller state each).  Returns dict(T=[B,n_save,Nz], naccept, nreject,
    nfe (RHS evaluations), status)."""
    dt_ = dtype
    T0 = np.array(T0, dtype=dt_)
    B, Nz = T0.shape
    theta = np.asarray(theta, dtype=dt_)
    colp = np.asarray(colp, dtype=dt_)
    saveat = np.asarray(saveat, dtype=dt_)
    n_save = saveat.size
    tf = dt_(saveat[-1])
    reltol, abstol = dt_(reltol), dt_(abstol)
    tab = rkc_coefficients().astype(dt_)
    mmax = rkc_max_stages(reltol, dt_)
    uround = dt_(np.finfo(dt_).eps)
    sqrtu = dt_(np.sqrt(uround))
    hmax, hmin = tf, dt_(10) * uround * tf
    third = dt_(1.0 / 3.0)
    out = np.zeros((B, n_save, Nz), dtype=dt_)
    naccept, nreject, nfev = np.zeros(B, np.int64), np.zeros(B, np.int64), np.zeros(B, np.int64)
    status = np.zeros(B, np.int32)

    def nrm(x):
        return dt_(np.sqrt(np.sum(x * x, dtype=dt_)))

    def rms(x):
        return dt_(np.sqrt(np.sum(x * x, dtype=dt_) * dt_(1.0 / Nz)))

    for col in range(B):
        nfe = 0

        def f(y):
            nonlocal nfe
            nfe += 1
            return rhs(y[None], theta, conv_width, layers, colp[col:col + 1], glob, dt_)[0]

        yn = T0[col].copy()
        si = 0
        while si < n_save and saveat[si] <= 0:
            out[col, si] = yn
            si += 1
        fn = f(yn)
        v = fn.copy()
        t = dt_(0)
        newspc, jacatt, first = True, False, True
        nstsig = nacc = nrej = 0
        absh = hmax
        errold = hlast = dt_(0)
        sprad = dt_(0)
        st = 0
        while t < tf:
            if nacc + nrej >= maxiters:
                st = 1
                break
            if newspc:
                # --- nonlinear power iteration on the Jacobian of f at y_n (RKC `rho`) ---
                ynrm, vnrm = nrm(yn), nrm(v)
                if ynrm != 0 and vnrm != 0:
                    dynrm = ynrm * sqrtu
                    v = yn + v * (dynrm / vnrm)
                elif ynrm != 0:
                    dynrm = ynrm * sqrtu
                    v = yn + yn * sqrtu
                elif vnrm != 0:
                    dynrm = uround
                    v = v * (dynrm / vnrm)
                else:
                    dynrm = uround
                    v = np.full(Nz, dynrm / dt_(np.sqrt(Nz)), dtype=dt_)
                # The RHS is only piecewise smooth (relu, min(0, K_CA dT/dz)): the iterate flips sign every round (J is
                # diffusion-like) and with it the active set, so sigma can settle into a 2-cycle.  Work with the upper envelope
                # max(sigma_k, sigma_{k-1}) -- the larger one-sided growth rate is the one stability needs -- and after 50
                # rounds take the largest value seen instead of giving up (RKC proper returns an error there).
                sigma = sigmal = env = smax = dt_(0)
                for it in range(1, 51):
                    fv = f(v)
                    dfnrm = nrm(fv - fn)
                    sigmall, sigmal, sigma = sigmal, sigma, dfnrm / dynrm
                    envl, env = env, max(sigma, sigmal)
                    smax = max(smax, sigma)
                    if it >= 3 and abs(env - envl) <= max(env, dt_(1) / hmax) * dt_(0.01):
                        smax = env
                        break
                    if dfnrm != 0 and it < 50:
                        v = yn + (fv - fn) * (dynrm / dfnrm)
                v = v - yn
                if not np.isfinite(smax):
                    st = 2
                    break
                sprad = dt_(1.2) * smax
                jacatt = True
            if first:
                absh = hmax
                if sprad * absh > 1:
                    absh = dt_(1) / sprad
                absh = max(absh, hmin)
                f1 = f(yn + absh * fn)
                est = absh * rms((f1 - fn) / (abstol + reltol * np.abs(yn)))
                if dt_(0.1) * absh < hmax * np.sqrt(est):
                    absh = max(dt_(0.1) * absh / dt_(np.sqrt(est)), hmin)
                else:
                    absh = hmax
                first = False
            last = False
            if dt_(1.1) * absh >= tf - t:
                absh, last = tf - t, True
            m = 1 + int(np.sqrt(dt_(1.54) * absh * sprad + dt_(1)))
            if m > mmax:
                m = mmax
                absh, last = dt_(m * m - 1) / (dt_(1.54) * sprad), False
            h = dt_(absh)
            c = tab[m]
            yjm2, yjm1 = yn, yn + (h * c[1, 3]) * fn
            for j in range(2, m + 1):
                fj = f(yjm1)
                y = c[j, 0] * yjm1 + c[j, 1] * yjm2 + c[j, 2] * yn + h * (c[j, 3] * fj + c[j, 4] * fn)
                yjm2, yjm1 = yjm1, y
            y = yjm1
            fy = f(y)
            est = dt_(0.8) * (yn - y) + (dt_(0.4) * h) * (fn + fy)
            err = rms(est / (abstol + reltol * np.maximum(np.abs(y), np.abs(yn))))
            if not np.isfinite(err):
                st = 2
                break
            if err > 1:
                nrej += 1
                absh = dt_(0.8) * absh / dt_(np.power(err, third))
                if absh < hmin:
                    st = 3
                    break
                newspc = not jacatt
                continue
            nacc += 1
            tn = tf if last else t + h
            d = (y - yn) / h
            c2, c3 = dt_(3) * d - dt_(2) * fn - fy, dt_(-2) * d + fn + fy
            while si < n_save and saveat[si] <= tn:       # cubic Hermite on (y_n, f_n, y_{n+1}, f_{n+1})
                th = min((saveat[si] - t) / h, dt_(1))
                out[col, si] = yn + (h * th) * (fn + th * (c2 + th * c3))
                si += 1
            jacatt = False
            nstsig = (nstsig + 1) % 25
            newspc = nstsig == 0
            fac = dt_(10)
            if nacc == 1:
                t2 = dt_(np.power(err, third))
                if dt_(0.8) < fac * t2:
                    fac = dt_(0.8) / t2
            else:
                t1 = dt_(0.8) * absh * dt_(np.power(errold, third))
                t2 = hlast * dt_(np.power(err, dt_(2) * third))
                if t1 < fac * t2:
                    fac = t1 / t2
            absh = max(dt_(0.1), fac) * absh
            absh = max(hmin, min(hmax, absh))
            errold, hlast = err, h
            yn, fn, t = y, fy, tn
        naccept[col], nreject[col], nfev[col], status[col] = nacc, nrej, nfe, st
    return dict(T=out, naccept=naccept, nreject=nreject, nfe=nfev, status=status)


def truth_solve(T0, theta, conv_width, layers, colp, glob, saveat, rtol=1e-12, atol=1e-12):
    """Tight-tolerance float64 solution of the same ODE (scipy DOP853) to bound everyone's error."""
    from scipy.integrate import solve_ivp
    T0 = np.asarray(T0, dtype=np.float64)
    B, Nz = T0.shape
    saveat = np.asarray(saveat, dtype=np.float64)
    out = np.zeros((B, saveat.size, Nz))
    for b in range(B):
        cp = np.asarray(colp[b:b + 1], dtype=np.float64)
        sol = solve_ivp(lambda t, y: rhs(y[None], theta, conv_width, layers, cp, glob)[0],
                        (saveat[0], saveat[-1]), T0[b], method="DOP853", t_eval=saveat, rtol=rtol, atol=atol)
        out[b] = sol.y.T
    return out


# ----------------------------------------------------------------------------------
# Loss (src/training.jl:45-48, Flux.Losses.mse = mean((yhat-y)^2)) and its gradient
# ----------------------------------------------------------------------------------
def mse(Tsol, Ttrue):
    d = np.asarray(Tsol, dtype=np.float64) - np.asarray(Ttrue, dtype=np.float64)
    return float(np.mean(d * d))


def loss_and_grad_autograd(T0, theta, conv_width, layers, colp, glob, saveat, Ttrue, tapes=None,
                           reltol=RELTOL_DEFAULT, abstol=ABSTOL_DEFAULT, fixed_dt=0.0):
    """float64 torch.autograd gradient of mse(solve(theta), Ttrue) w.r.t. theta through the accepted
    Tsit5 step sequence (step sizes held constant, as any discretise-then-differentiate adjoint does).
    Stands in for Zygote + DiffEqSensitivity (src/training.jl:70) -- see DESIGN.md on continuous vs
    discrete adjoints.  Returns (loss, grad[n_params], dL/dT0)."""
    import torch
    f64 = np.float64
    if tapes is None:
        tapes = tsit5_solve(T0, theta, conv_width, layers, colp, glob, saveat, reltol, abstol, f64,
                            fixed_dt=fixed_dt, record_tape=True)["tape"]
    th = torch.tensor(np.asarray(theta, dtype=f64), requires_grad=True)
    A, R = torch.tensor(TSIT5_A), torch.tensor(TSIT5_R)
    sT, swT, H, tau = [float(x) for x in glob]
    pref = swT / sT * tau / H
    B, Nz = np.asarray(T0).shape
    saveat = np.asarray(saveat, dtype=f64)
    n_save = saveat.size
    Ttrue_t = torch.tensor(np.asarray(Ttrue, dtype=f64))

    def nn_t(x):
        off = 0
        h = x
        if conv_width:
            w = th[off:off + conv_width]
            b = th[off + conv_width]
            off += conv_width + 1
            n0 = Nz - conv_width + 1
            acc = sum(w[j] * h[(conv_width - 1 - j):(conv_width - 1 - j) + n0] for j in range(conv_width))
            h = torch.relu(acc + b)
        for n_in, n_out, act in layers:
            W = th[off:off + n_in * n_out].reshape(n_in, n_out).T
            off += n_in * n_out
            b = th[off:off + n_out]
            off += n_out
            h = W @ h + b
            if act == ACT_RELU:
                h = torch.relu(h)
        return h

    def f_t(T, cp):
        F = torch.cat([cp[0:1], nn_t(T), cp[1:2]])
        dTdz = (T[1:] - T[:-1]) * Nz
        F = F - torch.cat([torch.zeros(1, dtype=T.dtype), torch.clamp(cp[2] * dTdz, max=0.0),
                           torch.zeros(1, dtype=T.dtype)])
        return -pref * (F[1:] - F[:-1]) * Nz

    total = torch.zeros((), dtype=torch.float64)
    T0_t = torch.tensor(np.asarray(T0, dtype=f64), requires_grad=True)
    for b in range(B):
        cp = torch.tensor(np.asarray(colp[b], dtype=f64))
        u = T0_t[b]
        si = 0
        while si < n_save and saveat[si] <= 0.0:
            total = total + ((u - Ttrue_t[b, si]) ** 2).sum()
            si += 1
        k1 = f_t(u, cp)
        for (t, h, _) in tapes[b]:
            ks = [k1]
            for s in range(1, 7):
                y = u + h * sum(A[s, j] * ks[j] for j in range(s))
                ks.append(f_t(y, cp))
                if s == 6:
                    unew = y
            tn = t + h
            if abs(tn - saveat[-1]) < 1e-9:
                tn = saveat[-1]
            while si < n_save and saveat[si] <= tn:
                thv = min((saveat[si] - t) / h, 1.0)
                bth = R @ torch.tensor([thv, thv ** 2, thv ** 3, thv ** 4], dtype=torch.float64)
                val = u + h * sum(bth[j] * ks[j] for j in range(7))
                total = total + ((val - Ttrue_t[b, si]) ** 2).sum()
                si += 1
            u, k1 = unew, ks[6]
        assert si == n_save, (si, n_save)
    loss = total / (B * n_save * Nz)
    loss.backward()
    return float(loss.detach()), th.grad.numpy().copy(), T0_t.grad.numpy().copy()


# ----------------------------------------------------------------------------------
# embedded path, src/oceananigans_nn.jl:3-30 (convective_adjustment!) and :158-282
# ----------------------------------------------------------------------------------
def convective_adjustment_matrix(T, dt, dz, K, bottom_gradient=0.0):
    """(ld, d, ud) of convective_adjustment! (src/oceananigans_nn.jl:10-24) for one column T[Nz] (index 0 = bottom).
    kappa_i = K where the cell-centre dT/dz < 0; the centre gradient is the mean of the two face gradients, the bottom
    face carries the gradient BC of base_model (:84), the top (flux BC) face is zero gradient [UPSTREAM halo filling]."""
    T = np.asarray(T)
    N = T.size
    gface = np.empty(N + 1, dtype=T.dtype)
    gface[1:N] = (T[1:] - T[:-1]) / dz
    gface[0], gface[N] = bottom_gradient, 0
    kap = np.where(0.5 * (gface[:-1] + gface[1:]) < 0, K, 0).astype(T.dtype)
    a = dt / dz**2
    ld = -a * kap[1:]                                   # ld[i] couples row i+1 to i  (Julia: ld = [-a kappa[i] for i in 2:Nz])
    ud = -a * kap[1:]                                   # ud[i] couples row i to i+1  (Julia: ud = [-a kappa[i+1] for i in 1:Nz-1])
    d = np.empty(N, dtype=T.dtype)
    d[:-1] = 1 + a * (kap[:-1] + kap[1:])
    d[-1] = 1 + a * kap[-1]
    return ld, d, ud


def thomas(ld, d, ud, rhs):
    """Tridiagonal solve (what `Tridiagonal(ld, d, ud) \\ T` does, src/oceananigans_nn.jl:26-27)."""
    n = d.size
    c, r = np.array(ud, dtype=d.dtype), np.array(rhs, dtype=d.dtype)
    b = np.array(d)
    for i in range(1, n):
        w = ld[i - 1] / b[i - 1]
        b[i] -= w * c[i - 1]
        r[i] -= w * r[i - 1]
    x = np.empty(n, dtype=d.dtype)
    x[-1] = r[-1] / b[-1]
    for i in range(n - 2, -1, -1):
        x[i] = (r[i] - c[i] * x[i + 1]) / b[i]
    return x


def embedded_solve(T0, heat_flux, theta, conv_width, layers, scal, Lz, dt, K, bottom_gradient, n_steps, save_every=1,
                   dtype=np.float64):
    """Batched restatement of oceananigans_convective_adjustment_with_neural_network (src/oceananigans_nn.jl:158-282):
    forward-Euler NN flux step (stand-in for Oceananigans' own time stepper [UPSTREAM]) + implicit convective adjustment.
    scal = (mu_T, sigma_T, mu_wT, sigma_wT).  Returns T[B, n_steps//save_every + 1, Nz] (physical units)."""
    T = np.array(T0, dtype=dtype)
    B, Nz = T.shape
    muT, sT, muw, sw = [dtype(x) for x in scal]
    dz, dt, K = dtype(Lz / Nz), dtype(dt), dtype(K)
    Q = np.asarray(heat_flux, dtype=dtype)
    out = np.empty((B, n_steps // save_every + 1, Nz), dtype=dtype)
    out[:, 0] = T
    th = np.asarray(theta, dtype=dtype)
    for step in range(1, n_steps + 1):
        F = np.zeros((B, Nz + 1), dtype=dtype)
        F[:, 1:Nz] = sw * nn_forward(th, conv_width, layers, (T - muT) / sT) + muw
        F[:, Nz] = Q
        T = T - dt * (F[:, 1:] - F[:, :-1]) / dz
        for b in range(B):
            ld, d, ud = convective_adjustment_matrix(T[b], dt, dz, K, dtype(bottom_gradient))
            T[b] = thomas(ld, d, ud, T[b])
        if step % save_every == 0:
            out[:, step // save_every] = T
    return out
