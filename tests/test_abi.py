"""CPU tests of the boundary: the C-ABI library loads, exports every symbol include/nfc.h declares, the ctypes
struct mirrors the C struct, and the host-side mirror of the reference API behaves (no compute calls: no GPU here)."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    hdr = open(os.path.join(ROOT, "include", "nfc.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(nfc_[a-z0-9_]+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol(nfc):
    lib = nfc.load()
    names = _declared()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/nfc.h but not exported by libnfc_b200.so"
    assert set(nfc.EXPORTS) == set(names)


def test_config_struct_matches_c_layout(nfc, tmp_path):
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "nfc.h"\nint main(){printf("%zu %zu %zu %zu %zu\\n",'
                   'sizeof(nfc_config), offsetof(nfc_config, saveat), offsetof(nfc_config, device), sizeof(nfc_counters),'
                   'offsetof(nfc_counters, last_kernel_ms));return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.check_call(["/usr/bin/gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    sz, off_save, off_dev, szc, off_ms = map(int, subprocess.check_output([str(exe)]).split())
    B = nfc._binding
    assert C.sizeof(B.NfcConfig) == sz and B.NfcConfig.saveat.offset == off_save and B.NfcConfig.device.offset == off_dev
    assert C.sizeof(B.NfcCounters) == szc and B.NfcCounters.last_kernel_ms.offset == off_ms


def test_no_cpu_fallback(nfc):
    """Without a GPU nfc_create must fail loudly (never compute on the CPU)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from oracle import nde_oracle as O
    with pytest.raises(nfc.NfcError, match="no CUDA device|CUDA"):
        nfc.Handle(0, O.chain_spec()[1], 1, np.linspace(0, 1, 9))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "neuralfreeconvection.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                uses = re.search(r"^\s*(import oracle|from oracle)|#include\s+\".*oracle|libnde_oracle|c_oracle", txt, flags=re.M)
                assert not uses, f"{f} uses the oracle"


def test_destructure_roundtrip_and_layout(nfc):
    from oracle import nde_oracle as O
    for arch in ["dense-default", "dense-wider", "dense-deeper", "conv-2", "conv-4"]:
        NN = nfc.named_chain(arch, seed=1)
        theta, re_ = nfc.destructure(NN)
        cw, layers = O.chain_spec(arch)
        assert theta.size == O.n_params(cw, layers) and NN.conv_width == cw and NN.dense_spec == layers
        conv, dense = O.unpack_theta(theta, cw, layers)
        dl = [l for l in NN.layers if isinstance(l, nfc.Dense)]
        for (W, b, _), l in zip(dense, dl):
            assert np.array_equal(W, l.weight) and np.array_equal(b, l.bias)
        NN2 = re_(theta * 2)
        assert np.array_equal(NN2.layers[-1].weight, NN.layers[-1].weight * 2)
        assert np.array_equal(nfc.destructure(NN2)[0], theta * 2)


def test_parameter_vectors_follow_reference_order(nfc):
    syn = nfc.synthetic
    Ts, wTs = syn.reference_scalings()
    zc = syn.cell_centres()
    zf = -256.0 + np.arange(33) * 8.0
    ds = nfc.ColumnDataset(T=np.zeros((32, 1153)), wT=np.zeros((33, 1153)), zc=zc, zf=zf, times=np.arange(1153) * 600.0,
                           metadata={"temperature_flux": 5.1e-6})
    p6 = nfc.FreeConvectionNDEParameters(ds, Ts, wTs)
    p7 = nfc.ConvectiveAdjustmentNDEParameters(ds, Ts, wTs, 2)
    assert p6.dtype == np.float32 and p6.shape == (6,) and p7.shape == (7,) and p7[6] == 2
    np.testing.assert_allclose(p6, [wTs(0.0), wTs(5.1e-6), Ts.sigma, wTs.sigma, 256.0, 691200.0], rtol=1e-6)
    np.testing.assert_array_equal(p7[:6], p6)


def test_scaling_is_zero_mean_unit_variance(nfc):
    x = np.random.default_rng(0).normal(3.0, 2.0, (32, 100))
    s = nfc.ZeroMeanUnitVarianceScaling.fit(x)
    y = s(x)
    assert abs(y.mean()) < 1e-12 and abs(y.std(ddof=1) - 1) < 1e-12
    np.testing.assert_allclose(s.inv(y), x, rtol=1e-12)


def test_adam_matches_flux_update_rule(nfc):
    opt = nfc.ADAM()
    th = np.ones(4)
    g = np.array([1.0, -2.0, 0.5, 0.0])
    th1 = opt.update(th, g)
    # first step of ADAM moves every non-zero-gradient coordinate by ~eta
    np.testing.assert_allclose(th1, th - 1e-3 * np.sign(g) * (np.abs(g) / (np.abs(g) + 1e-8)), rtol=1e-6)
    # several steps with changing gradients against Flux's rule written out (Flux.Optimise.ADAM: eta 1e-3, beta (0.9, 0.999), eps 1e-8;
    # m = b1 m + (1 - b1) g, v = b2 v + (1 - b2) g^2, theta -= eta * m / (1 - b1^t) / (sqrt(v / (1 - b2^t)) + eps))
    opt, th = nfc.ADAM(), np.array([0.3, -1.2, 2.0, 0.0])
    ref, m, v = th.copy(), np.zeros(4), np.zeros(4)
    rng = np.random.default_rng(0)
    for t in range(1, 8):
        g = rng.standard_normal(4) * np.array([1.0, 10.0, 0.01, 0.0])
        th = opt.update(th, g)
        m = 0.9 * m + 0.1 * g
        v = 0.999 * v + 0.001 * g * g
        ref = ref - 1e-3 * (m / (1 - 0.9 ** t)) / (np.sqrt(v / (1 - 0.999 ** t)) + 1e-8)
        np.testing.assert_allclose(th, ref, rtol=1e-12, atol=1e-15)


def test_on_disk_formats_round_trip(tmp_path):
    """the training driver's files carry the reference's keys (src/training.jl:3-12, train_free_convection_nde.jl:259-266,
    :376-390) as plain arrays in .npz; julia/convert_formats.jl rebuilds the JLD2 files from them."""
    import nfc_b200
    F, syn = nfc_b200.formats, nfc_b200.synthetic
    NN = nfc_b200.named_chain("conv-2", seed=3)
    theta, _ = nfc_b200.destructure(NN)
    Ts, wTs = syn.Scaling(19.6, 0.5), syn.Scaling(0.0, 1e-5)
    hist = str(tmp_path / "history")
    for e in (1, 2, 3):
        F.inscribe_history(hist, e, theta * e, 0.5 / e, 0.1 * e, NN)
    h = F.read_history(hist)
    assert list(h["epochs"]) == [1, 2, 3] and np.allclose(h["mean_loss"], [0.5, 0.25, 0.5 / 3]) and np.array_equal(h["neural_network"][2], theta * 3)
    assert list(h["arch"]) == [2, 31, 128, 128, 31]
    assert np.array_equal(nfc_b200.load_history(hist), h["neural_network"])           # what compute_nde_solution_history reads
    raw = dict(np.load(hist + ".npz"))
    assert {k.split("/")[0] for k in raw} == set(F.HISTORY_KEYS) | {"arch"}
    net = str(tmp_path / "neural_network_trained_on_timeseries")
    F.save_neural_network(net, NN, Ts, wTs)
    assert set(k.split("/")[0] for k in np.load(net + ".npz")) == set(F.NETWORK_FILE_KEYS)
    NN2, Ts2, wTs2, Nz = F.load_neural_network(net)
    assert Nz == 32 and (Ts2.mu, Ts2.sigma, wTs2.sigma) == (19.6, 0.5, 1e-5) and np.array_equal(nfc_b200.destructure(NN2)[0], theta)
    assert NN2.conv_width == 2 and [s[:2] for s in NN2.dense_spec] == [s[:2] for s in NN.dense_spec]
    sol = str(tmp_path / "solutions_and_history")
    T = np.arange(32 * 5, dtype=np.float64).reshape(32, 5)
    F.save_solutions(sol, NN, Ts, wTs, {"true": {1: {"T": T, "wT": T[:, :4]}}, "nde": {1: {"T": T + 1}, 2: {"T": T + 2}}},
                     solution_history={1: np.zeros((3, 32, 5))}, solution_loss_history={1: np.ones((3, 5))})
    s = F.load_solutions(sol)
    assert np.array_equal(s["nde"][2]["T"], T + 2) and np.array_equal(s["true"][1]["wT"], T[:, :4]) and s["solution_loss_history"][1].shape == (3, 5)
    with pytest.raises(ValueError):
        F.save_solutions(sol, NN, Ts, wTs, {"les": {}})


def _c_prototypes():
    """{name: [C parameter types]} of every function include/nfc.h declares."""
    import re
    src = open(os.path.join(ROOT, "include", "nfc.h")).read()
    src = re.sub(r"/\*.*?\*/", " ", src, flags=re.S)
    protos = {}
    for m in re.finditer(r"\b(?:int32_t|int64_t|const char\*)\s+(nfc_\w+)\s*\(([^;]*?)\)\s*;", src, flags=re.S):
        params = [p.strip() for p in m.group(2).replace("\n", " ").split(",") if p.strip() and p.strip() != "void"]
        protos[m.group(1)] = [re.sub(r"\s+\w+$", "", p).replace("const ", "").strip() for p in params]      # drop the parameter name
    return protos


def _julia_ccalls():
    """[(name, [Julia argument types])] of every ccall in julia/FreeConvectionB200.jl."""
    import re
    src = open(os.path.join(ROOT, "julia", "FreeConvectionB200.jl")).read()
    out = []
    for m in re.finditer(r"ccall\(\(:(nfc_\w+), LIB\),\s*\w+,\s*\(", src):
        i, depth = m.end(), 1
        while depth:                                   # the type tuple, with nested braces / parentheses
            depth += src[i] in "({"
            depth -= src[i] in ")}"
            i += 1
        tup = src[m.end():i - 1]
        types, cur, d = [], "", 0
        for ch in tup:
            if ch == "," and d == 0:
                types.append(cur.strip()); cur = ""
            else:
                d += ch in "({"; d -= ch in ")}"; cur += ch
        if cur.strip():
            types.append(cur.strip())
        out.append((m.group(1), types))
    return out


def test_julia_ccalls_match_the_header():
    """The Julia host cannot run in this image: at least every ccall in julia/FreeConvectionB200.jl must name a function of include/nfc.h
    with the same number of arguments and ABI-compatible types (pointer <-> Ptr / Ref, int32_t <-> Int32, int64_t <-> Int64, float <-> Float32)."""
    protos = _c_prototypes()
    calls = _julia_ccalls()
    assert len(calls) >= 12 and len(protos) >= 25
    kind = lambda c: "ptr" if c.endswith("*") else {"int32_t": "Int32", "int64_t": "Int64", "float": "Float32", "double": "Float64"}[c]
    for name, jt in calls:
        assert name in protos, f"{name} is not declared in include/nfc.h"
        ct = protos[name]
        assert len(ct) == len(jt), (name, ct, jt)
        for c, j in zip(ct, jt):
            k = kind(c)
            if k == "ptr":
                assert j.startswith(("Ptr{", "Ref{")) or j == "Cstring", (name, c, j)
            else:
                assert j == k, (name, c, j)
    # the struct the shim passes to nfc_create has the header's field order
    import re
    jl = open(os.path.join(ROOT, "julia", "FreeConvectionB200.jl")).read()
    body = re.search(r"struct NfcConfig\n(.*?)\nend", jl, flags=re.S).group(1)
    jfields = [f.split("::")[0].strip() for l in body.split("\n") for f in l.split(";") if "::" in f]
    hdr = re.sub(r"/\*.*?\*/", " ", open(os.path.join(ROOT, "include", "nfc.h")).read(), flags=re.S)
    cbody = re.search(r"typedef struct nfc_config \{(.*?)\} nfc_config;", hdr, flags=re.S).group(1)
    cfields = [re.sub(r"\[.*?\]", "", f).strip().split()[-1].lstrip("*") for f in cbody.split(";") if f.strip()]
    assert jfields == cfields, (jfields, cfields)
