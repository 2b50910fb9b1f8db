"""On-disk formats of the training driver.

The reference writes JLD2 files whose `neural_network` entries are serialised Julia structs (`Flux.Chain`): such a file can only be
produced by Julia itself.  This module therefore writes the SAME KEYS with plain arrays into `.npz` containers (readable from Julia with
NPZ.jl), and `julia/convert_formats.jl` turns them into exactly the JLD2 files the reference's `figure*.jl` / `movie*.jl` scripts open:

  history file (src/training.jl:3-12, :62)                neural_network/<e>, mean_loss/<e>, runtime/<e>
  neural_network_trained_on_timeseries                     grid_points, neural_network, T_scaling, wT_scaling
      (train_free_convection_nde.jl:259-266)
  solutions_and_history (train_free_convection_nde.jl:376-390, plus the histories of src/testing.jl:42-79 that the figures read)
                                                           grid_points, neural_network, T_scaling, wT_scaling,
                                                           true / nde / kpp / tke / convective_adjustment / oceananigans: id => solution,
                                                           solution_history, solution_loss_history: id => array

Representation of the Julia objects inside the npz:
  neural_network   `<key>/theta` Float32[n_params] in Flux.destructure order + `<key>/arch` = [conv_width, in_0, out_0, ..., out_L]
                   (hidden activations relu, last identity: the five Chains of train_free_convection_nde.jl:125-156)
  *_scaling        `<key>` = Float64[2] = (mu, sigma) of ZeroMeanUnitVarianceScaling
  id => solution   `<group>/<id>/<field>` for every field of the solution (T, wT, ...)
"""
from __future__ import annotations

import os

import numpy as np

HISTORY_KEYS = ("neural_network", "mean_loss", "runtime")
NETWORK_FILE_KEYS = ("grid_points", "neural_network", "T_scaling", "wT_scaling")
SOLUTION_GROUPS = ("true", "nde", "kpp", "tke", "convective_adjustment", "oceananigans")
HISTORY_GROUPS = ("solution_history", "solution_loss_history")


def _npz(path):
    return path if path.endswith(".npz") else path + ".npz"


def chain_arch(NN):
    """[conv_width, in_0, out_0, out_1, ..., out_L] of a Chain of the reference's five architectures."""
    dims = [NN.dense_spec[0][0]] + [spec[1] for spec in NN.dense_spec]
    return np.asarray([NN.conv_width] + dims, dtype=np.int64)


def chain_from_arch(arch):
    """A Chain (fresh weights) of the architecture chain_arch() describes."""
    from .api import Chain, Conv, Dense, identity, relu
    arch = [int(x) for x in arch]
    cw, dims = arch[0], arch[1:]
    layers = [Conv(cw, relu)] if cw else []
    layers += [Dense(dims[i], dims[i + 1], relu if i < len(dims) - 2 else identity) for i in range(len(dims) - 1)]
    return Chain(*layers)


def _put_network(d, key, NN, theta=None):
    from .api import destructure
    d[f"{key}/theta"] = np.asarray(destructure(NN)[0] if theta is None else theta, dtype=np.float32)
    d[f"{key}/arch"] = chain_arch(NN)


def _put_scaling(d, key, s):
    d[key] = np.asarray([s.mu, s.sigma], dtype=np.float64)


def inscribe_history(path, epoch, theta, loss, runtime, NN=None):
    """inscribe_history (src/training.jl:3-12): append neural_network/<e>, mean_loss/<e>, runtime/<e>."""
    if path is None:
        return
    path = _npz(path)
    prev = dict(np.load(path)) if os.path.exists(path) else {}
    prev[f"neural_network/{epoch}"] = np.asarray(theta, dtype=np.float32)
    prev[f"mean_loss/{epoch}"] = np.float64(loss)
    prev[f"runtime/{epoch}"] = np.float64(runtime)
    if NN is not None and "arch" not in prev:
        prev["arch"] = chain_arch(NN)
    np.savez(path, **prev)


def read_history(path):
    """-> dict(epochs, neural_network [E][n_params], mean_loss [E], runtime [E], arch or None)"""
    f = dict(np.load(_npz(path)))
    epochs = sorted(int(k.split("/")[1]) for k in f if k.startswith("neural_network/"))
    return dict(epochs=np.asarray(epochs), neural_network=np.stack([f[f"neural_network/{e}"] for e in epochs]),
                mean_loss=np.asarray([float(f[f"mean_loss/{e}"]) for e in epochs]),
                runtime=np.asarray([float(f[f"runtime/{e}"]) for e in epochs]), arch=f.get("arch"))


def save_neural_network(path, NN, T_scaling, wT_scaling, grid_points=32):
    """neural_network_trained_on_timeseries (train_free_convection_nde.jl:259-266)."""
    d = {"grid_points": np.int64(grid_points)}
    _put_network(d, "neural_network", NN)
    _put_scaling(d, "T_scaling", T_scaling)
    _put_scaling(d, "wT_scaling", wT_scaling)
    np.savez(_npz(path), **d)


def load_neural_network(path):
    """-> (NN, T_scaling, wT_scaling, grid_points) from a file written by save_neural_network."""
    from .api import destructure
    from .synthetic import Scaling
    f = dict(np.load(_npz(path)))
    NN = chain_from_arch(f["neural_network/arch"])
    _, re = destructure(NN)
    NN = re(np.asarray(f["neural_network/theta"], dtype=np.float32))
    sc = lambda k: Scaling(float(f[k][0]), float(f[k][1]))
    return NN, sc("T_scaling"), sc("wT_scaling"), int(f["grid_points"])


def save_solutions(path, NN, T_scaling, wT_scaling, solutions, solution_history=None, solution_loss_history=None, grid_points=32):
    """solutions_and_history (train_free_convection_nde.jl:376-390).  `solutions`: {"true" | "nde" | "kpp" | "tke" |
    "convective_adjustment" | "oceananigans": {simulation id: {field: array}}}; the two histories: {simulation id: array}."""
    d = {"grid_points": np.int64(grid_points)}
    _put_network(d, "neural_network", NN)
    _put_scaling(d, "T_scaling", T_scaling)
    _put_scaling(d, "wT_scaling", wT_scaling)
    for group, sols in solutions.items():
        if group not in SOLUTION_GROUPS:
            raise ValueError(f"unknown solution group '{group}' (the reference writes {SOLUTION_GROUPS})")
        for sim, fields in sols.items():
            for name, arr in (fields.items() if isinstance(fields, dict) else {"T": fields}.items()):
                d[f"{group}/{sim}/{name}"] = np.asarray(arr)
    for group, h in (("solution_history", solution_history), ("solution_loss_history", solution_loss_history)):
        for sim, arr in (h or {}).items():
            d[f"{group}/{sim}"] = np.asarray(arr)
    np.savez(_npz(path), **d)


def load_solutions(path):
    """-> nested dict mirroring what save_solutions was given (plus grid_points, scalings, the network's theta / arch)."""
    f = dict(np.load(_npz(path)))
    out = {"grid_points": int(f["grid_points"]), "T_scaling": f["T_scaling"], "wT_scaling": f["wT_scaling"],
           "neural_network": {"theta": f["neural_network/theta"], "arch": f["neural_network/arch"]}}
    for k, v in f.items():
        parts = k.split("/")
        if parts[0] in SOLUTION_GROUPS and len(parts) == 3:
            out.setdefault(parts[0], {}).setdefault(int(parts[1]) if parts[1].lstrip("-").isdigit() else parts[1], {})[parts[2]] = v
        elif parts[0] in HISTORY_GROUPS and len(parts) == 2:
            out.setdefault(parts[0], {})[int(parts[1]) if parts[1].lstrip("-").isdigit() else parts[1]] = v
    return out
