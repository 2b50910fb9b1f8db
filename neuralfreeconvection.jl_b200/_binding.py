"""ctypes binding of libnfc_b200.so (include/nfc.h).  No CPU fallback: if the CUDA library is missing or
no B200 is visible, construction raises."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("NFC_B200_LIB") or os.path.join(_HERE, "libnfc_b200.so")   # NFC_B200_LIB: a debug build of the same library (build.sh)
NFC_MAX_LAYERS = 8

EXPORTS = ["nfc_create", "nfc_destroy", "nfc_last_error", "nfc_n_params", "nfc_set_weights", "nfc_set_saveat",
           "nfc_get_counters", "nfc_rhs", "nfc_solve", "nfc_diagnose_flux", "nfc_loss_grad", "nfc_rhs_dev",
           "nfc_solve_dev", "nfc_loss_grad_dev", "nfc_sync", "nfc_measure_fp32_peak", "nfc_get_adjoint_steps", "nfc_embedded_solve",
           "nfc_embedded_diagnose_flux", "nfc_solve_ensemble", "nfc_comm_unique_id", "nfc_comm_init", "nfc_comm_destroy", "nfc_get_tape", "nfc_set_option"]


class NfcConfig(C.Structure):
    _fields_ = [("struct_size", C.c_int32), ("Nz", C.c_int32), ("conv_width", C.c_int32), ("n_layers", C.c_int32),
                ("layer_dims", C.c_int32 * (NFC_MAX_LAYERS + 1)), ("activations", C.c_int32 * NFC_MAX_LAYERS),
                ("nde_type", C.c_int32), ("alg", C.c_int32), ("reltol", C.c_float), ("abstol", C.c_float),
                ("fixed_dt", C.c_float), ("max_iters", C.c_int32), ("n_save", C.c_int32),
                ("saveat", C.POINTER(C.c_float)), ("device", C.c_int32), ("adjoint_max_steps", C.c_int32),
                ("adjoint_reltol", C.c_float), ("adjoint_abstol", C.c_float), ("reserved", C.c_int32 * 8)]


class NfcCounters(C.Structure):
    _fields_ = [("columns", C.c_int64), ("steps_accepted", C.c_int64), ("steps_rejected", C.c_int64),
                ("adj_steps_accepted", C.c_int64), ("adj_steps_rejected", C.c_int64), ("kernel_launches", C.c_int64),
                ("last_kernel_ms", C.c_float), ("last_adjoint_ms", C.c_float), ("rhs_evals", C.c_int64),
                ("last_allreduce_ms", C.c_float), ("comm_world", C.c_int32)]


_lib = None


def load():
    """dlopen the in-tree CUDA library.  Raises (never falls back) when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing: run ./build.sh (nvcc, sm_100a). There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    fp, ip, dp, vp = C.POINTER(C.c_float), C.POINTER(C.c_int32), C.POINTER(C.c_double), C.c_void_p
    lib.nfc_create.argtypes = [C.POINTER(NfcConfig), C.POINTER(vp)]
    lib.nfc_destroy.argtypes = [vp]
    lib.nfc_last_error.argtypes = [vp]
    lib.nfc_last_error.restype = C.c_char_p
    lib.nfc_n_params.argtypes = [vp]
    lib.nfc_n_params.restype = C.c_int64
    lib.nfc_set_weights.argtypes = [vp, fp, C.c_int64]
    lib.nfc_set_saveat.argtypes = [vp, fp, C.c_int32]
    lib.nfc_get_counters.argtypes = [vp, C.POINTER(NfcCounters)]
    lib.nfc_rhs.argtypes = [vp, fp, fp, fp, fp, C.c_int64]
    lib.nfc_solve.argtypes = [vp, fp, fp, fp, fp, ip, ip, ip, C.c_int64]
    lib.nfc_diagnose_flux.argtypes = [vp, fp, fp, fp, C.c_int64]
    lib.nfc_solve_ensemble.argtypes = [vp, fp, C.c_int64, fp, fp, fp, fp, ip, ip, ip, C.c_int64]
    lib.nfc_loss_grad.argtypes = [vp, fp, fp, fp, fp, dp, fp, ip, C.c_int64]
    lib.nfc_rhs_dev.argtypes = [vp, vp, vp, vp, vp, C.c_int64, vp]
    lib.nfc_solve_dev.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, C.c_int64, vp]
    lib.nfc_loss_grad_dev.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, C.c_int64, C.c_int64, vp]
    lib.nfc_sync.argtypes = [vp]
    lib.nfc_embedded_solve.argtypes = [vp, fp, fp, fp, C.c_float, C.c_float, C.c_float, C.c_float, C.c_int32, C.c_int32, fp, C.c_int64]
    lib.nfc_embedded_diagnose_flux.argtypes = [vp, fp, fp, fp, C.c_float, C.c_float, C.c_int32, fp, C.c_int64]
    lib.nfc_embedded_diagnose_flux.restype = C.c_int32
    lib.nfc_get_adjoint_steps.argtypes = [vp, ip, C.c_int64]
    lib.nfc_measure_fp32_peak.argtypes = [C.c_int32, C.c_float, dp]
    lib.nfc_get_tape.argtypes = [vp, C.c_int64, fp, C.c_int32, ip]
    lib.nfc_set_option.argtypes = [vp, C.c_char_p, C.c_int64]
    lib.nfc_comm_unique_id.argtypes = [vp]
    lib.nfc_comm_init.argtypes = [vp, vp, C.c_int32, C.c_int32]
    lib.nfc_comm_destroy.argtypes = [vp]
    for name in EXPORTS:
        if name not in ("nfc_last_error", "nfc_n_params"):
            getattr(lib, name).restype = C.c_int32
    _lib = lib
    return lib


def _f32(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int32)) if a is not None else None


class NfcError(RuntimeError):
    pass


def measure_fp32_peak(device=0, ms_target=50.0):
    """Sustained FP32 FMA TFLOP/s of `device` (pure-FFMA kernel inside libnfc_b200)."""
    out = C.c_double(0.0)
    rc = load().nfc_measure_fp32_peak(device, ms_target, C.byref(out))
    if rc != 0:
        raise NfcError(f"nfc_measure_fp32_peak failed ({rc}): {load().nfc_last_error(None).decode()}")
    return out.value


NFC_UNIQUE_ID_BYTES = 128


def comm_unique_id() -> bytes:
    """nfc_comm_unique_id: rank 0 creates the NCCL rendezvous id; the host ships these 128 bytes to the other ranks."""
    buf = C.create_string_buffer(NFC_UNIQUE_ID_BYTES)
    rc = load().nfc_comm_unique_id(buf)
    if rc != 0:
        raise NfcError(f"nfc_comm_unique_id failed ({rc}): {load().nfc_last_error(None).decode()}")
    return buf.raw


class Handle:
    """Owns one nfc_handle (device buffers, stream).  Thin, 1:1 with include/nfc.h."""

    def __init__(self, conv_width, layers, nde_type, saveat, reltol=1e-4, abstol=1e-6, fixed_dt=0.0, device=0,
                 max_iters=100000, adjoint_max_steps=0, adjoint_reltol=0.0, adjoint_abstol=0.0, Nz=32, alg=0):
        self.lib = load()
        cfg = NfcConfig()
        cfg.struct_size = C.sizeof(NfcConfig)
        cfg.Nz = Nz
        cfg.conv_width = conv_width
        if len(layers) > NFC_MAX_LAYERS:
            raise NfcError("too many Dense layers")
        cfg.n_layers = len(layers)
        dims = [layers[0][0]] + [o for _, o, _ in layers]
        for i, d in enumerate(dims):
            cfg.layer_dims[i] = d
        for i, (_, _, a) in enumerate(layers):
            cfg.activations[i] = a
        cfg.nde_type = nde_type
        cfg.alg = {"Tsit5": 0, "RKC2": 1, "ROCK4": 1}.get(alg, alg) if isinstance(alg, str) else int(alg)
        cfg.reltol, cfg.abstol, cfg.fixed_dt = reltol, abstol, fixed_dt
        cfg.max_iters = max_iters
        self.saveat = _f32(saveat)
        cfg.n_save = self.saveat.size
        cfg.saveat = _fp(self.saveat)
        cfg.device = device
        cfg.adjoint_max_steps = adjoint_max_steps
        cfg.adjoint_reltol, cfg.adjoint_abstol = adjoint_reltol, adjoint_abstol
        self.Nz = Nz
        self.n_save = int(cfg.n_save)
        self.device = device
        h = C.c_void_p()
        rc = self.lib.nfc_create(C.byref(cfg), C.byref(h))
        if rc != 0:
            raise NfcError(f"nfc_create failed ({rc}): {self.lib.nfc_last_error(None).decode()}")
        self.h = h
        self.comm_world = 1
        self.n_params = int(self.lib.nfc_n_params(h))

    def _check(self, rc, what):
        if rc != 0:
            raise NfcError(f"{what} failed ({rc}): {self.lib.nfc_last_error(self.h).decode()}")

    def close(self):
        if getattr(self, "h", None):
            self.lib.nfc_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_weights(self, theta):
        theta = _f32(theta)
        self._check(self.lib.nfc_set_weights(self.h, _fp(theta), theta.size), "nfc_set_weights")

    def set_saveat(self, saveat):
        self.saveat = _f32(saveat)
        self.n_save = self.saveat.size
        self._check(self.lib.nfc_set_saveat(self.h, _fp(self.saveat), self.n_save), "nfc_set_saveat")

    def counters(self):
        c = NfcCounters()
        self._check(self.lib.nfc_get_counters(self.h, C.byref(c)), "nfc_get_counters")
        return {k: getattr(c, k) for k, _ in NfcCounters._fields_}

    # ---- host-buffer calls ----
    def rhs(self, T, colp, glob):
        T, colp, glob = _f32(T), _f32(colp), _f32(glob)
        B = T.shape[0]
        out = np.empty_like(T)
        self._check(self.lib.nfc_rhs(self.h, _fp(T), _fp(colp), _fp(glob), _fp(out), B), "nfc_rhs")
        return out

    def solve(self, T0, colp, glob, out=None):
        T0, colp, glob = _f32(T0), _f32(colp), _f32(glob)
        B = T0.shape[0]
        if out is None:
            out = np.empty((B, self.n_save, self.Nz), dtype=np.float32)
        na, nr, st = (np.zeros(B, dtype=np.int32) for _ in range(3))
        self._check(self.lib.nfc_solve(self.h, _fp(T0), _fp(colp), _fp(glob), _fp(out), _ip(na), _ip(nr), _ip(st), B),
                    "nfc_solve")
        return dict(T=out, naccept=na, nreject=nr, status=st)

    def solve_ensemble(self, theta_members, T0, colp, glob):
        """nfc_solve_ensemble: member e integrates T0[e] ([E][S][Nz]) with its own weights theta_members[e] (src/testing.jl:42-79)."""
        theta_members, T0, colp, glob = _f32(theta_members), _f32(T0), _f32(colp), _f32(glob)
        if theta_members.ndim != 2 or theta_members.shape[1] != self.n_params:
            raise ValueError(f"theta_members must be [E][{self.n_params}], got {theta_members.shape}")
        E = theta_members.shape[0]
        if T0.ndim != 3 or T0.shape[0] != E or colp.shape[:2] != T0.shape[:2]:
            raise ValueError("T0 must be [E][S][Nz] and colp [E][S][3]")
        S = T0.shape[1]
        out = np.empty((E, S, self.n_save, self.Nz), dtype=np.float32)
        na, nr, st = (np.zeros((E, S), dtype=np.int32) for _ in range(3))
        self._check(self.lib.nfc_solve_ensemble(self.h, _fp(theta_members), E, _fp(T0), _fp(colp), _fp(glob), _fp(out), _ip(na), _ip(nr),
                                                _ip(st), S), "nfc_solve_ensemble")
        return dict(T=out, naccept=na, nreject=nr, status=st)

    def diagnose_flux(self, Tsol, colp):
        Tsol, colp = _f32(Tsol), _f32(colp)
        B = Tsol.shape[0]
        wT = np.empty((B, self.n_save, self.Nz + 1), dtype=np.float32)
        self._check(self.lib.nfc_diagnose_flux(self.h, _fp(Tsol), _fp(colp), _fp(wT), B), "nfc_diagnose_flux")
        return wT

    def loss_grad(self, T0, colp, glob, Ttrue):
        T0, colp, glob, Ttrue = _f32(T0), _f32(colp), _f32(glob), _f32(Ttrue)
        B = T0.shape[0]
        grad = np.zeros(self.n_params, dtype=np.float32)
        st = np.zeros(max(B, 1), dtype=np.int32)[:B]
        loss = C.c_double(0.0)
        self._check(self.lib.nfc_loss_grad(self.h, _fp(T0), _fp(colp), _fp(glob), _fp(Ttrue), C.byref(loss), _fp(grad),
                                           _ip(st), B), "nfc_loss_grad")
        return dict(loss=loss.value, grad=grad, status=st)

    # ---- device-buffer calls (torch tensors on this handle's device; plumbing only) ----
    @staticmethod
    def _glob(glob):
        g = _f32(glob.detach().cpu().numpy() if hasattr(glob, "detach") else glob)
        return g, g.ctypes.data

    def rhs_dev(self, T, colp, glob, out, stream=None):
        g, gp = self._glob(glob)
        self._check(self.lib.nfc_rhs_dev(self.h, T.data_ptr(), colp.data_ptr(), gp, out.data_ptr(),
                                         T.shape[0], stream), "nfc_rhs_dev")

    def solve_dev(self, T0, colp, glob, Tout, naccept=None, nreject=None, status=None, stream=None):
        p = lambda t: t.data_ptr() if t is not None else None
        g, gp = self._glob(glob)
        self._check(self.lib.nfc_solve_dev(self.h, T0.data_ptr(), colp.data_ptr(), gp, Tout.data_ptr(),
                                           p(naccept), p(nreject), p(status), T0.shape[0], stream), "nfc_solve_dev")

    def loss_grad_dev(self, T0, colp, glob, Ttrue, loss_sum, grad, status=None, n_total_columns=None, stream=None):
        B = T0.shape[0]
        g, gp = self._glob(glob)
        self._check(self.lib.nfc_loss_grad_dev(self.h, T0.data_ptr(), colp.data_ptr(), gp, Ttrue.data_ptr(),
                                               loss_sum.data_ptr(), grad.data_ptr(),
                                               status.data_ptr() if status is not None else None, B,
                                               n_total_columns if n_total_columns is not None else B, stream),
                    "nfc_loss_grad_dev")

    def embedded_solve(self, T0, heat_flux, scal, Lz, dt, K, bottom_gradient, n_steps, save_every=1):
        """nfc_embedded_solve: explicit NN flux step + implicit convective adjustment (tridiagonal solve) per fixed step."""
        T0, heat_flux, scal = _f32(T0), _f32(heat_flux), _f32(scal)
        B = T0.shape[0]
        out = np.empty((B, n_steps // save_every + 1, self.Nz), dtype=np.float32)
        self._check(self.lib.nfc_embedded_solve(self.h, _fp(T0), _fp(heat_flux), _fp(scal), Lz, dt, K, bottom_gradient, n_steps, save_every,
                                                _fp(out), B), "nfc_embedded_solve")
        return out

    def adjoint_steps(self, n):
        out = np.zeros(n, dtype=np.int32)
        self._check(self.lib.nfc_get_adjoint_steps(self.h, _ip(out), n), "nfc_get_adjoint_steps")
        return out

    def sync(self):
        self._check(self.lib.nfc_sync(self.h), "nfc_sync")

    def embedded_diagnose_flux(self, Tsol, heat_flux, scal, Lz, K):
        """nfc_embedded_diagnose_flux: wT [B][n_out][Nz+1] (physical units) of saved physical profiles Tsol [B][n_out][Nz]."""
        Tsol, heat_flux, scal = _f32(Tsol), _f32(heat_flux), _f32(scal)
        B, n_out = Tsol.shape[0], Tsol.shape[1]
        wT = np.empty((B, n_out, self.Nz + 1), dtype=np.float32)
        self._check(self.lib.nfc_embedded_diagnose_flux(self.h, _fp(Tsol), _fp(heat_flux), _fp(scal), Lz, K, n_out,
                                                        _fp(wT), B), "nfc_embedded_diagnose_flux")
        return wT

    def set_option(self, name, value):
        """nfc_set_option: kernel selection / scratch budgets (include/nfc.h lists the names)."""
        self._check(self.lib.nfc_set_option(self.h, name.encode(), int(value)), "nfc_set_option")

    def tape(self, column, max_rows=4096):
        """nfc_get_tape: [n_rows][168] forward tape of one column of the last loss_grad chunk (u, c1..c4, t, h per accepted step)."""
        rows = np.zeros((max_rows, 168), dtype=np.float32)
        n = C.c_int32(0)
        self._check(self.lib.nfc_get_tape(self.h, column, _fp(rows), max_rows, C.byref(n)), "nfc_get_tape")
        return rows[:min(n.value, max_rows)]

    def comm_init(self, unique_id: bytes, rank: int, world: int):
        """nfc_comm_init (collective): from now on loss_grad / loss_grad_dev return the loss and gradient of the GLOBAL batch."""
        if len(unique_id) != NFC_UNIQUE_ID_BYTES:
            raise ValueError("the rendezvous id has 128 bytes")
        buf = C.create_string_buffer(unique_id, NFC_UNIQUE_ID_BYTES)
        self._check(self.lib.nfc_comm_init(self.h, buf, rank, world), "nfc_comm_init")
        self.comm_world = world

    def comm_destroy(self):
        self._check(self.lib.nfc_comm_destroy(self.h), "nfc_comm_destroy")
        self.comm_world = 1
