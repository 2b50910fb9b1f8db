"""Host-side mirror of the reference's API for the NDE hot path (Julia is not available in this image, so
the host language is Python over the C ABI; `julia/FreeConvectionB200.jl` is the ccall shim a maintainer
would drop into the reference).  Names, argument order and error behaviour follow the reference:

    Chain / Dense / Conv / destructure          Flux 0.12.6 as used at train_free_convection_nde.jl:125-156, src/solve.jl:2
    ZeroMeanUnitVarianceScaling                 train_free_convection_nde.jl:216-217 (OceanParameterizations)
    FreeConvectionNDE, ...Parameters            src/free_convection_nde.jl:1-62
    ConvectiveAdjustmentNDE, ...Parameters      src/convective_adjustment_nde.jl:1-72
    solve_nde (2 methods)                       src/solve.jl:1-51
    train_neural_differential_equation          src/training.jl:34-78 (Julia: train_neural_differential_equation!)
    ADAM                                        Flux.ADAM defaults (train_free_convection_nde.jl:245)

All numerics run in libnfc_b200.so on the GPU; nothing here evaluates the network or integrates on the CPU.
"""
from __future__ import annotations

import time
from dataclasses import dataclass, field

from typing import NamedTuple

import numpy as np

from . import _binding as B_
from .synthetic import Scaling as ZeroMeanUnitVarianceScaling  # same object as the synthetic generator uses

ACT_IDENTITY, ACT_RELU = 0, 1
identity, relu = ACT_IDENTITY, ACT_RELU


# ------------------------------------------------------------------------------------------------
# Flux-like containers (parameters only; evaluation happens on the GPU)
# ------------------------------------------------------------------------------------------------
class Dense:
    """Dense(in, out, act): weight is out x in, bias out (Flux field names `weight`, `bias`)."""

    def __init__(self, n_in, n_out, act=identity, rng=None):
        rng = rng or np.random.default_rng()
        lim = np.sqrt(6.0 / (n_in + n_out))                       # Flux glorot_uniform
        self.weight = rng.uniform(-lim, lim, (n_out, n_in)).astype(np.float32)
        self.bias = np.zeros(n_out, dtype=np.float32)
        self.act = act


class Conv:
    """Conv((k,1), 1=>1, relu) front-end of the conv-2 / conv-4 Chains (train_free_convection_nde.jl:138-153)."""

    def __init__(self, k, act=relu, rng=None):
        rng = rng or np.random.default_rng()
        lim = np.sqrt(6.0 / (k + k))
        self.weight = rng.uniform(-lim, lim, k).astype(np.float32)
        self.bias = np.zeros(1, dtype=np.float32)
        self.act = act


class Chain:
    def __init__(self, *layers):
        self.layers = list(layers)

    def params(self):
        return [p for l in self.layers for p in (l.weight, l.bias)]

    @property
    def conv_width(self):
        return self.layers[0].weight.size if isinstance(self.layers[0], Conv) else 0

    @property
    def dense_spec(self):
        return [(l.weight.shape[1], l.weight.shape[0], l.act) for l in self.layers if isinstance(l, Dense)]


def named_chain(arch="dense-default", Nz=32, seed=0):
    """The five architectures of train_free_convection_nde.jl:125-153."""
    rng = np.random.Generator(np.random.PCG64(seed))
    if arch == "dense-default":
        return Chain(Dense(Nz, 4 * Nz, relu, rng), Dense(4 * Nz, 4 * Nz, relu, rng), Dense(4 * Nz, Nz - 1, identity, rng))
    if arch == "dense-wider":
        return Chain(Dense(Nz, 8 * Nz, relu, rng), Dense(8 * Nz, 8 * Nz, relu, rng), Dense(8 * Nz, Nz - 1, identity, rng))
    if arch == "dense-deeper":
        return Chain(Dense(Nz, 4 * Nz, relu, rng), Dense(4 * Nz, 4 * Nz, relu, rng), Dense(4 * Nz, 4 * Nz, relu, rng),
                     Dense(4 * Nz, Nz - 1, identity, rng))
    if arch in ("conv-2", "conv-4"):
        k = int(arch[-1])
        return Chain(Conv(k, relu, rng), Dense(Nz - k + 1, 4 * Nz, relu, rng), Dense(4 * Nz, 4 * Nz, relu, rng),
                     Dense(4 * Nz, Nz - 1, identity, rng))
    raise ValueError(f"Invalid neural network architecture: {arch}")


def destructure(NN: Chain):
    """Flux.destructure: flat Float32 vector (layer order; weight column-major, then bias) and a re-builder."""
    theta = np.concatenate([np.asarray(p, dtype=np.float32).reshape(-1, order="F") for p in NN.params()])

    def reconstruct(flat):
        flat = np.asarray(flat, dtype=np.float32)
        out = Chain(*[_clone(l) for l in NN.layers])
        off = 0
        for l in out.layers:
            for name in ("weight", "bias"):
                p = getattr(l, name)
                setattr(l, name, flat[off:off + p.size].reshape(p.shape, order="F").copy())
                off += p.size
        return out                     # like Flux's _restructure, extra trailing elements are ignored

    return theta, reconstruct


def _clone(l):
    c = object.__new__(type(l))
    c.weight, c.bias, c.act = l.weight.copy(), l.bias.copy(), l.act
    return c


# ------------------------------------------------------------------------------------------------
# Datasets: the slice of Oceananigans' FieldDataset the NDE path touches (ds["T"], ds["wT"], metadata)
# ------------------------------------------------------------------------------------------------
@dataclass
class ColumnDataset:
    T: np.ndarray                 # [Nz, Nt]   horizontally averaged temperature (unscaled)
    wT: np.ndarray                # [Nz+1, Nt] turbulent heat flux on faces (unscaled)
    zc: np.ndarray                # [Nz] cell centres (negative, index 0 deepest)
    zf: np.ndarray                # [Nz+1]
    times: np.ndarray             # [Nt] seconds
    metadata: dict = field(default_factory=dict)   # "temperature_flux"


class _NDE:
    """What `FreeConvectionNDE(NN, ds; iterations)` returns in the reference: an ODEProblem with tspan/saveat
    (src/free_convection_nde.jl:40-46); here it also owns the GPU handle."""

    nde_type = None

    def __init__(self, NN, ds, iterations=None, device=0, reltol=1e-4, abstol=1e-6, alg=0, **kw):
        Nz, Nt = ds.T.shape
        self.Nz, self.Nt = Nz, Nt
        self.H = abs(float(ds.zf[0]))
        self.tau = float(ds.times[-1])
        if iterations is None:
            iterations = range(1, Nt + 1)
        iterations = np.asarray(list(iterations))
        self.iterations = iterations
        self.tspan = (np.float32(0.0), np.float32(iterations.max() / Nt))
        self.saveat = np.linspace(self.tspan[0], self.tspan[1], len(iterations)).astype(np.float32)
        self._mk = lambda a, **kw2: B_.Handle(NN.conv_width, NN.dense_spec, self.nde_type, self.saveat, reltol=reltol, abstol=abstol,
                                              device=device, Nz=Nz, alg=a, **{**kw, **kw2})
        self._handles = {}
        self.handle = self.handle_for(alg)

    def handle_for(self, alg):
        """The solver is fixed per C handle; the reference passes it per solve call (src/solve.jl:1), so keep one handle per algorithm."""
        a = alg if isinstance(alg, int) else _check_alg(alg)
        if a not in self._handles:
            self._handles[a] = self._mk(a)
        return self._handles[a]


class FreeConvectionNDE(_NDE):
    nde_type = 0


class ConvectiveAdjustmentNDE(_NDE):
    nde_type = 1


def FreeConvectionNDEParameters(ds, T_scaling, wT_scaling):
    """[bottom_flux, top_flux, sigma_T, sigma_wT, H, tau] as Float32 (src/free_convection_nde.jl:49-62)."""
    H, tau = abs(float(ds.zf[0])), float(ds.times[-1])
    return np.array([wT_scaling(ds.wT[0, 0]), wT_scaling(ds.metadata["temperature_flux"]), T_scaling.sigma,
                     wT_scaling.sigma, H, tau], dtype=np.float32)


def ConvectiveAdjustmentNDEParameters(ds, T_scaling, wT_scaling, K_CA):
    """7-vector with trailing K_CA (src/convective_adjustment_nde.jl:59-72)."""
    return np.concatenate([FreeConvectionNDEParameters(ds, T_scaling, wT_scaling), np.float32([K_CA])])


def _split_params(nde, nde_params):
    p = np.asarray(nde_params, dtype=np.float32)
    want = 6 if nde.nde_type == 0 else 7
    if p.shape[-1] != want:       # the reference would mis-slice silently; be loud instead
        raise ValueError(f"{type(nde).__name__} needs {want} trailing parameters, got {p.shape[-1]}")
    p2 = np.atleast_2d(p)
    colp = np.zeros((p2.shape[0], 3), dtype=np.float32)
    colp[:, 0:2] = p2[:, 0:2]
    if want == 7:
        colp[:, 2] = p2[:, 6]
    glob = p2[0, 2:6].copy()
    if not np.all(p2[:, 2:6] == glob):
        raise ValueError("sigma_T, sigma_wT, H, tau must be shared by all columns of one call")
    return colp, glob


def solve_nde(*args, **kw):
    """solve_nde(nde, NN, T0, alg, nde_params) -> Array [Nz, n_save]  (src/solve.jl:1-6), or batched: T0 [B, Nz]
    with nde_params [B, 6|7] -> [B, n_save, Nz];
    solve_nde(ds, NN, NDEType, nde_params, algorithm, T_scaling, wT_scaling; T0=None) -> (T, wT)  (src/solve.jl:8-51)."""
    if isinstance(args[0], _NDE):
        nde, NN, T0, alg, nde_params = args
        handle = nde.handle_for(alg)
        theta, _ = destructure(NN)
        handle.set_weights(theta)
        colp, glob = _split_params(nde, nde_params)
        T0 = np.asarray(T0, dtype=np.float32)
        single = T0.ndim == 1
        res = handle.solve(np.atleast_2d(T0), colp, glob)
        nde.last = res
        _warn_status(res["status"], "solve_nde")
        return res["T"][0].T.copy() if single else res["T"]
    ds, NN, NDEType, nde_params, algorithm, T_scaling, wT_scaling = args
    T0 = kw.get("T0")
    nde = NDEType(NN, ds)
    if T0 is None:
        T0 = T_scaling(ds.T[:, 0])
    if len(nde_params) < 7:
        raise IndexError("solve_nde(ds, ...) reads nde_params[7] (src/solve.jl:30): pass ConvectiveAdjustmentNDEParameters")
    p = np.asarray(nde_params, dtype=np.float32)
    T = solve_nde(nde, NN, T0, algorithm, p if nde.nde_type == 1 else p[:6])          # [Nz, Nt]
    colp, _ = _split_params(nde, p if nde.nde_type == 1 else p[:6])
    wT = nde.handle_for(algorithm).diagnose_flux(T.T[None], colp)[0].T                  # [Nz+1, Nt]; the handle that holds the weights
    return dict(T=T_scaling.inv(T), wT=wT_scaling.inv(wT))


STATUS_NAMES = {1: "maxiters", 2: "NaN", 3: "dt underflow", 4: "adjoint tape overflow", 5: "FP16x2 operand range"}


def _warn_status(status, what):
    """The reference treats solver retcodes as warnings, never exceptions: so does the host side here -- but it says so.
    Failed columns return NaN after the failure (nfc_solve) and contribute nothing to the gradient (nfc_loss_grad)."""
    status = np.asarray(status)
    if status.any():
        import warnings
        kinds = {STATUS_NAMES.get(int(k), str(k)): int((status == k).sum()) for k in np.unique(status) if k != 0}
        warnings.warn(f"{what}: {int((status != 0).sum())} of {status.size} columns failed ({kinds}); their results are NaN / excluded", RuntimeWarning, stacklevel=3)


def _check_alg(alg):
    """Algorithm id of the C ABI: Tsit5 -> NFC_ALG_TSIT5; RKC2 / ROCK4 -> NFC_ALG_RKC2 (stabilised explicit, see class ROCK4)."""
    name = (alg if isinstance(alg, str) else type(alg).__name__).replace("()", "")
    if name == "Tsit5":
        return 0
    if name in ("RKC2", "ROCK4"):
        return 1
    raise NotImplementedError(f"time stepper {name}: this build implements Tsit5 and the stabilised explicit RKC2 (ROCK4's role)")


class Tsit5:
    pass


class RKC2:
    """Second-order Runge-Kutta-Chebyshev (Sommeijer, Shampine & Verwer 1998): explicit, stage count chosen per step from a
    spectral-radius estimate, so the convective-adjustment term does not limit the step as it does for Tsit5."""


class ROCK4(RKC2):
    """The reference's production integrator (train_free_convection_nde.sh:3).  ROCK4's coefficient tables are OrdinaryDiffEq's
    and not available offline; selecting it here runs RKC2, the same kind of method at lower order: trajectories agree with
    any convergent integrator to the tolerance, step and evaluation counts differ from ROCK4's."""


class ADAM:
    """Flux.ADAM(eta=1e-3, beta=(0.9, 0.999)), eps = 1e-8 [UPSTREAM Flux 0.12.6]."""

    def __init__(self, eta=1e-3, beta=(0.9, 0.999), eps=1e-8):
        self.eta, self.beta, self.eps = eta, beta, eps
        self.m = self.v = None
        self.bp = [1.0, 1.0]

    def update(self, theta, g):
        if self.m is None:
            self.m, self.v = np.zeros_like(theta), np.zeros_like(theta)
        b1, b2 = self.beta
        self.m = b1 * self.m + (1 - b1) * g
        self.v = b2 * self.v + (1 - b2) * g * g
        self.bp[0] *= b1
        self.bp[1] *= b2
        return theta - self.eta * (self.m / (1 - self.bp[0])) / (np.sqrt(self.v / (1 - self.bp[1])) + self.eps)


def train_neural_differential_equation(NN, NDEType, nde_params, algorithm, datasets, T_scaling, iterations, opt, epochs,
                                       history_filepath=None, device=0, callback=None):
    """train_neural_differential_equation!(...) (src/training.jl:34-78): full-batch ADAM on the MSE between the NDE
    solutions and the scaled truth of every training simulation.  One epoch = one nfc_loss_grad call (all simulations
    in one launch) + one ADAM update.  The loss recorded for an epoch is the loss of the weights its gradient was taken at (the
    reference's callback re-solves every simulation for that number, src/training.jl:53-54; nfc_loss_grad already has it).
    Returns the best-loss Chain (GalacticOptim returns sol.minimizer, src/training.jl:74)."""
    alg = _check_alg(algorithm)          # Tsit5, or the stabilised explicit integrator (ROCK4's role): the continuous adjoint reads either tape
    ids = sorted(datasets.keys())
    it = np.asarray(list(iterations)) - 1                                   # Julia 1-based -> 0-based
    T0 = np.stack([T_scaling(datasets[i].T[:, 0]) for i in ids]).astype(np.float32)
    ndes = {i: NDEType(NN, datasets[i], iterations=iterations, device=device) for i in ids[:1]}   # shared handle
    nde = ndes[ids[0]]
    true_sols = np.stack([T_scaling(datasets[i].T[:, it]).T for i in ids]).astype(np.float32)     # [B, n_save, Nz]
    P = np.stack([np.asarray(nde_params[i], dtype=np.float32) for i in ids])
    colp, glob = _split_params(nde, P if nde.nde_type == 1 else P[:, :6])
    theta, reconstruct = destructure(NN)
    best = (np.inf, theta.copy())
    history = []
    t0 = time.time()
    for epoch in range(1, epochs + 1):
        handle = nde.handle_for(alg)
        handle.set_weights(theta)
        res = handle.loss_grad(T0, colp, glob, true_sols)
        if (res["status"] == 4).any():
            # adjoint tape overflow (a column accepted more steps than the dense tape's adjoint_max_steps rows): its residuals are in the loss
            # but its gradient is missing -- never train on that.  The compact tape gives every column the rows it needs.
            handle.set_option("tape_mode", 2)
            res = handle.loss_grad(T0, colp, glob, true_sols)
        _warn_status(res["status"], f"train_neural_differential_equation epoch {epoch}")
        if res["loss"] < best[0]:
            best = (res["loss"], theta.copy())
        runtime = time.time() - t0
        history.append(dict(epoch=epoch, mean_loss=res["loss"], runtime=runtime))
        if callback is not None:
            callback(epoch, theta, res["loss"])
        _inscribe_history(history_filepath, epoch, theta, res["loss"], runtime, NN)
        t0 = time.time()
        theta = opt.update(theta, res["grad"]).astype(np.float32)
    NN_final = reconstruct(best[1])
    NN_final.history = history
    return NN_final


def _inscribe_history(path, epoch, theta, loss, runtime, NN=None):
    """inscribe_history (src/training.jl:3-12) -- keys neural_network/<e>, mean_loss/<e>, runtime/<e>; the container is .npz and
    `julia/convert_formats.jl` turns it into the reference's JLD2 file (formats.py)."""
    from . import formats
    formats.inscribe_history(path, epoch, theta, loss, runtime, NN)


# ------------------------------------------------------------------------------------------------
# epoch-history evaluators, src/testing.jl
# ------------------------------------------------------------------------------------------------
def load_history(history):
    """The `neural_network/<e>` entries of a training history (what inscribe_history wrote) as [epochs][n_params]."""
    if isinstance(history, str):
        history = dict(np.load(history if history.endswith(".npz") else history + ".npz"))
    if isinstance(history, dict):
        epochs = sorted(int(k.split("/")[1]) for k in history if k.startswith("neural_network/"))
        return np.stack([np.asarray(history[f"neural_network/{e}"], dtype=np.float32) for e in epochs])
    return np.asarray(history, dtype=np.float32)


def compute_nde_solution_history(datasets, NDEType, nde_params, algorithm, NN, history, T_scaling, device=0):
    """compute_nde_solution_history (src/testing.jl:42-79): the NDE solution of every simulation under every saved epoch's
    network.  The reference runs epochs x simulations solve_nde calls one after the other; here all of them are one
    nfc_solve_ensemble launch (member = epoch).  `NN` gives the architecture, `history` the per-epoch destructured weights
    (npz path / dict as written by inscribe_history, or an [epochs][n_params] array).
    Returns (solution_history, loss_history): {id: [epochs, Nz, Nt]} in physical units and {id: [epochs, Nt]} with
    Flux.Losses.mse(T_true, T, agg = x -> mean(x, dims=1)), i.e. the mean over depth per saved time."""
    if _check_alg(algorithm) != 0:
        raise NotImplementedError("the ensemble launch runs Tsit5")
    thetas = load_history(history)
    ids = sorted(datasets.keys())
    ds0 = datasets[ids[0]]
    nde = NDEType(NN, ds0, device=device)
    P = np.stack([np.asarray(nde_params[i], dtype=np.float32) for i in ids])
    if P.shape[-1] < 7:
        raise IndexError("solve_nde(ds, ...) reads nde_params[7] (src/solve.jl:30): pass ConvectiveAdjustmentNDEParameters")
    colp, glob = _split_params(nde, P if nde.nde_type == 1 else P[:, :6])
    T0 = np.stack([T_scaling(datasets[i].T[:, 0]) for i in ids]).astype(np.float32)          # [S, Nz]
    E = thetas.shape[0]
    res = nde.handle.solve_ensemble(thetas, np.broadcast_to(T0, (E,) + T0.shape), np.broadcast_to(colp, (E,) + colp.shape), glob)
    sol = T_scaling.inv(res["T"].astype(np.float64))                                           # [E, S, Nt, Nz]
    solution_history, loss_history = {}, {}
    for s_, i in enumerate(ids):
        solution_history[i] = np.transpose(sol[:, s_], (0, 2, 1))                              # [E, Nz, Nt]
        loss_history[i] = np.mean((datasets[i].T[None] - solution_history[i]) ** 2, axis=1)    # [E, Nt]
    nde.last = res
    return solution_history, loss_history


# ------------------------------------------------------------------------------------------------
# the embedded (Oceananigans-style) path, src/oceananigans_nn.jl
# ------------------------------------------------------------------------------------------------
def oceananigans_convective_adjustment_with_neural_network(ds, NN, T_scaling, wT_scaling, dt=600.0, K=10.0, device=0, save_every=1):
    """Fixed-step explicit NN-flux step + implicit convective adjustment (src/oceananigans_nn.jl:158-282).  Returns T [Nz, Nt]
    in physical units.  K is non-dimensionalised exactly like the reference does (:195)."""
    Nz, Nt = ds.T.shape
    Lz, stop_time = abs(float(ds.zf[0])), float(ds.times[-1])
    Knd = wT_scaling.sigma / T_scaling.sigma * stop_time / Lz * K
    return _embedded(ds, NN, T_scaling, wT_scaling, dt, Knd, device, save_every)


def oceananigans_convective_adjustment(ds, dt=600.0, K=10.0, device=0, save_every=1):
    """The pure convective-adjustment baseline (src/oceananigans_nn.jl:97-156): zero network flux, surface flux, implicit CA."""
    NN = named_chain("dense-default", Nz=ds.T.shape[0])
    for p_ in NN.params():
        p_ *= 0
    ident = ZeroMeanUnitVarianceScaling(0.0, 1.0)
    return _embedded(ds, NN, ident, ident, dt, K, device, save_every)


class EmbeddedSolution(NamedTuple):
    T: np.ndarray
    wT: np.ndarray


def _embedded(ds, NN, T_scaling, wT_scaling, dt, K, device, save_every):
    Nz, Nt = ds.T.shape
    Lz, stop_time = abs(float(ds.zf[0])), float(ds.times[-1])
    n_steps = int(round(stop_time / dt))
    h = B_.Handle(NN.conv_width, NN.dense_spec, 1, np.float32([0.0, 1.0]), device=device, Nz=Nz)
    theta, _ = destructure(NN)
    h.set_weights(theta)
    scal = [T_scaling.mu, T_scaling.sigma, wT_scaling.mu, wT_scaling.sigma]
    Q = [ds.metadata["temperature_flux"]]
    out = h.embedded_solve(ds.T[None, :, 0], Q, scal, Lz, dt, K,
                           float(ds.metadata.get("dθdz_deep", ds.metadata.get("dthetadz_deep", 0.0))), n_steps, save_every)
    wT = h.embedded_diagnose_flux(out, Q, scal, Lz, K)               # diagnose_wT_NN, src/oceananigans_nn.jl:200-218
    return EmbeddedSolution(T=out[0].T, wT=wT[0].T)                   # the reference's (T=T_nn, wT=wT_nn): Nz x Nt and (Nz+1) x Nt
