"""ctypes wrapper of oracle/nde_oracle.c (test / CPU-baseline infrastructure only)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force=False):
    so = os.path.join(_HERE, "libnde_oracle.so")
    src = os.path.join(_HERE, "nde_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-B", "-C", _HERE, "libnde_oracle.so"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
    return _LIB


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _p(a, t=C.c_float):
    return a.ctypes.data_as(C.POINTER(t)) if a is not None else None


def _net(conv_width, layers):
    dims = np.array([layers[0][0]] + [o for _, o, _ in layers], dtype=np.int32)
    acts = np.array([a for _, _, a in layers], dtype=np.int32)
    return C.c_int(conv_width), C.c_int(len(layers)), _p(dims, C.c_int), _p(acts, C.c_int), (dims, acts)


def rhs(T, theta, conv_width, layers, colp, glob, nthreads=1):
    T, theta, colp, glob = _f32(T), _f32(theta), _f32(colp), _f32(glob)
    B, Nz = T.shape
    out = np.empty_like(T)
    cw, nl, dims, acts, keep = _net(conv_width, layers)
    rc = lib().oracle_rhs(C.c_int(Nz), cw, nl, dims, acts, _p(theta), _p(T), _p(colp), _p(glob), _p(out),
                          C.c_int64(B), C.c_int(nthreads))
    assert rc == 0, rc
    return out


def solve(T0, theta, conv_width, layers, colp, glob, saveat, reltol=1e-4, abstol=1e-6, fixed_dt=0.0,
          nthreads=1, want_T=True):
    T0, theta, colp, glob, saveat = _f32(T0), _f32(theta), _f32(colp), _f32(glob), _f32(saveat)
    B, Nz = T0.shape
    n_save = saveat.size
    out = np.empty((B, n_save, Nz), dtype=np.float32) if want_T else None
    na = np.zeros(B, dtype=np.int32)
    nr = np.zeros(B, dtype=np.int32)
    st = np.zeros(B, dtype=np.int32)
    cw, nl, dims, acts, keep = _net(conv_width, layers)
    rc = lib().oracle_solve(C.c_int(Nz), cw, nl, dims, acts, _p(theta), _p(T0), _p(colp), _p(glob), _p(saveat),
                            C.c_int(n_save), C.c_float(reltol), C.c_float(abstol), C.c_float(fixed_dt), _p(out),
                            _p(na, C.c_int32), _p(nr, C.c_int32), _p(st, C.c_int32), C.c_int64(B), C.c_int(nthreads))
    assert rc == 0, rc
    return dict(T=out, naccept=na, nreject=nr, status=st)


def loss_grad(T0, theta, conv_width, layers, colp, glob, saveat, Ttrue, reltol=1e-4, abstol=1e-6, fixed_dt=0.0,
              nthreads=1):
    T0, theta, colp, glob, saveat, Ttrue = (_f32(a) for a in (T0, theta, colp, glob, saveat, Ttrue))
    B, Nz = T0.shape
    n_save = saveat.size
    grad = np.zeros(theta.size, dtype=np.float32)
    dT0 = np.zeros_like(T0)
    na = np.zeros(B, dtype=np.int32)
    loss = C.c_double(0.0)
    cw, nl, dims, acts, keep = _net(conv_width, layers)
    rc = lib().oracle_loss_grad(C.c_int(Nz), cw, nl, dims, acts, _p(theta), _p(T0), _p(colp), _p(glob), _p(saveat),
                                C.c_int(n_save), C.c_float(reltol), C.c_float(abstol), C.c_float(fixed_dt),
                                _p(Ttrue), C.byref(loss), _p(grad), _p(dT0), _p(na, C.c_int32), C.c_int64(B),
                                C.c_int(nthreads))
    assert rc == 0, rc
    return dict(loss=loss.value, grad=grad, dT0=dT0, naccept=na)


def loss_grad_continuous(T0, theta, conv_width, layers, colp, glob, saveat, Ttrue, reltol=1e-4, abstol=1e-6, fixed_dt=0.0,
                         adj_reltol=None, adj_abstol=None, nthreads=1):
    """Continuous (InterpolatingAdjoint-like) gradient: the algorithm the CUDA backward kernel implements."""
    T0, theta, colp, glob, saveat, Ttrue = (_f32(a) for a in (T0, theta, colp, glob, saveat, Ttrue))
    B, Nz = T0.shape
    n_save = saveat.size
    grad = np.zeros(theta.size, dtype=np.float32)
    dT0 = np.zeros_like(T0)
    na = np.zeros(B, dtype=np.int32)
    nr = np.zeros(B, dtype=np.int32)
    loss = C.c_double(0.0)
    cw, nl, dims, acts, keep = _net(conv_width, layers)
    rc = lib().oracle_loss_grad_continuous(C.c_int(Nz), cw, nl, dims, acts, _p(theta), _p(T0), _p(colp), _p(glob), _p(saveat),
                                           C.c_int(n_save), C.c_float(reltol), C.c_float(abstol), C.c_float(fixed_dt),
                                           C.c_float(adj_reltol if adj_reltol else reltol),
                                           C.c_float(adj_abstol if adj_abstol else abstol), _p(Ttrue), C.byref(loss), _p(grad),
                                           _p(dT0), _p(na, C.c_int32), _p(nr, C.c_int32), C.c_int64(B), C.c_int(nthreads))
    assert rc == 0, rc
    return dict(loss=loss.value, grad=grad, dT0=dT0, adj_naccept=na, adj_nreject=nr)
