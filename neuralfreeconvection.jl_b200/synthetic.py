"""Deterministic synthetic columns for the BASELINE configs.

The reference's inputs are horizontally averaged LES profiles fetched by DataDeps
(src/data.jl:7-12) -- unreachable offline -- so the harness synthesises columns from the LES
generator's own initial-condition recipe (three_layer_constant_fluxes.jl:242-267,347-361,
linear thermocline branch :324-325, no noise) and the run table run_les_simulations.sh:4-36.
Everything is NumPy; RNG = PCG64(seed).
"""
from __future__ import annotations

import numpy as np

ALPHA = 2e-4            # three_layer_constant_fluxes.jl:236  LinearEquationOfState(alpha=2e-4)
G = 9.80665             # Oceananigans gravitational_acceleration default [UPSTREAM]
H_DOMAIN = 256.0        # src/data.jl:36
TAU = 691200.0          # 192 h (run_les_simulations.sh:4), 10-minute output -> Nt = 1153
NT = 1153
THETA_SURFACE = 20.0

# run_les_simulations.sh:4-16 (the nine training simulations):
# (Qb, surface-layer depth, thermocline width, N2 surface, N2 thermocline, N2 deep)
TRAINING_RUNS = [
    (1e-8, 48, 24, 2e-6, 1e-5, 2e-6), (3e-8, 48, 24, 2e-6, 1e-5, 2e-6), (5e-8, 48, 24, 2e-6, 1e-5, 2e-6),
    (1e-8, 24, 48, 1e-6, 1.5e-5, 2e-6), (3e-8, 24, 48, 1e-6, 1.5e-5, 2e-6), (5e-8, 24, 48, 1e-6, 1.5e-5, 2e-6),
    (1e-8, 24, 64, 1e-6, 2e-5, 5e-6), (3e-8, 24, 64, 1e-6, 2e-5, 5e-6), (5e-8, 24, 64, 1e-6, 2e-5, 5e-6),
]


def cell_centres(Nz=32):
    dz = H_DOMAIN / Nz
    return -H_DOMAIN + (np.arange(Nz) + 0.5) * dz          # index 0 = deepest cell


def three_layer_profile(zc, sld, tw, N2s, N2t, N2d):
    """Piecewise-linear theta(z), vectorised over columns: zc[Nz], the rest [B]."""
    sld, tw, N2s, N2t, N2d = (np.asarray(a, dtype=np.float64)[:, None] for a in (sld, tw, N2s, N2t, N2d))
    z = zc[None, :]
    gs, gt, gd = N2s / (ALPHA * G), N2t / (ALPHA * G), N2d / (ALPHA * G)
    z_tr, z_dp = -sld, -(sld + tw)
    th_tr = THETA_SURFACE + z_tr * gs
    th_dp = th_tr + (z_dp - z_tr) * gt
    return np.where(z > z_tr, THETA_SURFACE + gs * z,
                    np.where(z > z_dp, th_tr + gt * (z - z_tr), th_dp + gd * (z - z_dp)))


def heat_flux(Qb):
    return np.asarray(Qb, dtype=np.float64) / (ALPHA * G)    # three_layer_constant_fluxes.jl:238


class Scaling:
    """ZeroMeanUnitVarianceScaling (OceanParameterizations@main [UPSTREAM]; fields mu/sigma
    src/oceananigans_nn.jl:180-181; callable src/solve.jl:24; inverse src/solve.jl:50)."""

    def __init__(self, mu, sigma):
        self.mu, self.sigma = float(mu), float(sigma)

    @classmethod
    def fit(cls, data):
        data = np.asarray(data, dtype=np.float64)
        return cls(data.mean(), data.std(ddof=1))            # Julia std = n-1 form

    def __call__(self, x):
        return (np.asarray(x, dtype=np.float64) - self.mu) / self.sigma

    def inv(self, y):
        return self.sigma * np.asarray(y, dtype=np.float64) + self.mu


def reference_scalings(Nz=32):
    """Fixed (T_scaling, wT_scaling) for all configs: mean/std over the nine training runs of the
    initial profiles and of linear-in-z flux profiles (train_free_convection_nde.jl:216-217 fits
    them to the whole training matrix; the LES time series is unavailable offline)."""
    zc = cell_centres(Nz)
    r = np.array(TRAINING_RUNS)
    T = three_layer_profile(zc, r[:, 1], r[:, 2], r[:, 3], r[:, 4], r[:, 5])
    zf = -H_DOMAIN + np.arange(Nz + 1) * (H_DOMAIN / Nz)
    wT = heat_flux(r[:, 0])[:, None] * (1.0 + zf[None, :] / H_DOMAIN)
    return Scaling.fit(T), Scaling.fit(wT)


def make_columns(kind, B=None, seed=1, Nz=32, K_CA=2.0):
    """Returns dict(T0[B,Nz] f32 scaled, colp[B,3] f32 (bottom, top, K_CA), glob[4] f32, T_scaling, wT_scaling).
    kind: 'single' (config 1), 'training9' (config 3), 'ocean' (configs 2 and 4: the envelope of the 21 runs)."""
    Ts, wTs = reference_scalings(Nz)
    zc = cell_centres(Nz)
    if kind == "single":
        r = np.array(TRAINING_RUNS[:1])
    elif kind == "training9":
        r = np.array(TRAINING_RUNS)
    elif kind == "ocean":
        rng = np.random.Generator(np.random.PCG64(seed))
        r = np.stack([rng.uniform(0.5e-8, 6e-8, B), rng.uniform(24, 48, B), rng.uniform(24, 64, B),
                      rng.choice([1e-6, 2e-6], B), rng.uniform(0.75e-5, 2.25e-5, B),
                      rng.choice([2e-6, 5e-6], B)], axis=1)
    else:
        raise ValueError(kind)
    n = r.shape[0]
    T0 = Ts(three_layer_profile(zc, r[:, 1], r[:, 2], r[:, 3], r[:, 4], r[:, 5]))
    colp = np.empty((n, 3))
    colp[:, 0] = wTs(0.0)                                    # FreeConvectionNDEParameters: bottom_flux
    colp[:, 1] = wTs(heat_flux(r[:, 0]))                     # top_flux = wT_scaling(temperature_flux)
    colp[:, 2] = K_CA
    glob = np.array([Ts.sigma, wTs.sigma, H_DOMAIN, TAU])    # (sigma_T, sigma_wT, H, tau)
    return dict(T0=T0.astype(np.float32), colp=colp.astype(np.float32), glob=glob.astype(np.float32),
                T_scaling=Ts, wT_scaling=wTs)


def saveat(n_save):
    """range(0, 1, length=n_save) as Float32 (src/free_convection_nde.jl:40-41 with iterations covering 1:Nt)."""
    return np.linspace(0.0, 1.0, n_save).astype(np.float32)
