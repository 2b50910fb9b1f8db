#!/usr/bin/env python
"""bench.py -- NDE column-steps/sec (forward solve) on N B200s of one node.

    python bench.py --gpus N --steps K --warmup W            (N>1: launched by torch.distributed.run, one rank per GPU)
    python bench.py --impl reference ...                     (the CPU port of the reference path on the host cores)

A "step" is one pass of the hot path over one batch: the batched Tsit5 solve of BASELINE config 2
(65,536 independent columns, varied surface flux / stratification, ConvectiveAdjustmentNDE K_CA = 2, default Chain,
129 save points = the training `iterations = 1:9:1153`), weights random-init as the reference (Glorot x 1e-5).
One column-step = one attempted Tsit5 step of one column (6 new RHS evaluations) -- SURVEY.md 8d.
Columns shard across ranks with no data-path collective (weak scaling: 65,536 columns per GPU).
Next to the headline the line carries BASELINE's second metric and config 4: `adjoint` (seconds per loss+gradient epoch) and `config4`
(1,048,576 columns sharded over the N ranks: forward solve, then loss + gradient with the NCCL all-reduce of [grad; loss] INSIDE the
timed region -- the collective runs inside nfc_loss_grad on the handle's own communicator, include/nfc.h nfc_comm_init).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "nde_column_steps_per_sec_fwd"
UNIT = "column-steps/s"
COLUMNS_PER_GPU = 65536
N_SAVE = 129
FLOPS_PER_COLUMN_STEP = 6 * 2 * (32 * 128 + 128 * 128 + 128 * 31)      # 293,376 (default Chain)
FLOPS_PER_ADJOINT_COLUMN_STEP = 3 * FLOPS_PER_COLUMN_STEP               # 880,128 (the kernel runs 7 stages: no FSAL backwards)
CONFIG4_COLUMNS = 1048576
WORKLOAD = ("config2: 65536 columns/GPU, ConvectiveAdjustmentNDE K_CA=2, dense-default Chain (24735 params, Glorot x1e-5), "
            "Tsit5 reltol 1e-4 abstol 1e-6, saveat 129 points on [0,1]")


def make_inputs(B, seed):
    import nfc_b200
    syn = nfc_b200.synthetic
    d = syn.make_columns("ocean", B=B, seed=seed, K_CA=2.0)
    NN = nfc_b200.named_chain("dense-default", seed=0)                  # Flux glorot_uniform, zero biases
    for p in NN.params():
        p *= 1e-5                                                       # train_free_convection_nde.jl:158-160
    theta, _ = nfc_b200.destructure(NN)
    return d, NN.conv_width, NN.dense_spec, theta, syn.saveat(N_SAVE)


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [x.strip() for x in line.split(",")]))

    def count(self, t0, t1):
        return sum(1 for ts, _ in self.rows if t0 <= ts <= t1)

    def stop(self, t0=None, t1=None, window="timed region"):
        """Clocks of the samples that arrived in [t0, t1] (host times; all samples without a window)."""
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        self.t.join(timeout=2)
        rows = [r for ts, r in self.rows if t0 is None or t0 <= ts <= t1 + 0.05]
        sm = [float(r[0]) for r in rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in rows if len(r) >= 7 for i in range(4) if r[3 + i].lower().startswith("active")})
        pw = [float(r[2]) for r in rows if len(r) > 2 and r[2].replace(".", "").isdigit()]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "window": window}


def bind_to_gpu_numa_node(index):
    """Pin this rank to the CPUs next to its GPU (the host-buffer path copies 1 GB per step through that root complex)."""
    try:
        import torch
        bus = torch.cuda.get_device_properties(index).pci_bus_id
        dom = getattr(torch.cuda.get_device_properties(index), "pci_domain_id", 0)
        path = f"/sys/bus/pci/devices/{dom:04x}:{bus:02x}:00.0/local_cpulist"
        cpus = set()
        for part in open(path).read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        if cpus:
            os.sched_setaffinity(0, cpus)
            return f"{min(cpus)}-{max(cpus)} ({len(cpus)} cpus, {path})"
    except Exception as e:      # best effort
        return f"unchanged ({type(e).__name__})"
    return "unchanged"


def cpu_reference_run(steps, warmup, sample_columns, nthreads):
    """Times the CPU port (oracle/nde_oracle.c, OpenMP over columns) on a bounded sample of the same workload."""
    from oracle import c_oracle as CO
    d, cw, layers, theta, sv = make_inputs(sample_columns, seed=1)
    times, nsteps = [], 0
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        r = CO.solve(d["T0"], theta, cw, layers, d["colp"], d["glob"], sv, nthreads=nthreads)
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
            nsteps = int((r["naccept"] + r["nreject"]).sum())
    return nsteps * len(times) / sum(times), 1e3 * sum(times) / len(times), nsteps


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1
    sample = 8192
    value, ms, nsteps = cpu_reference_run(args.steps, max(args.warmup, 1), sample, cores)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": {"workload": WORKLOAD, "note": "CPU port of the reference path; the Julia reference cannot run "
                                            "here (no julia in the image); port omits the reference's per-call allocations"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": f"{sample} of the 65536 columns per step, {nsteps} column-steps, OpenMP over columns"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line), flush=True)
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--columns", type=int, default=COLUMNS_PER_GPU, help="columns per GPU (default: BASELINE config 2)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-large-batch", action="store_true", help="skip the 1,048,576-column forward solve (config 4) reported next to the headline")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import nfc_b200

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the B200 path has no CPU fallback (use --impl reference for the CPU port)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    warmup = max(args.warmup, 3)
    B = args.columns
    affinity = bind_to_gpu_numa_node(local_rank)

    d, cw, layers, theta, sv = make_inputs(B, seed=1 + rank)
    h = nfc_b200.Handle(cw, layers, 1, sv, device=local_rank)
    h.set_weights(theta)
    T0, colp = (torch.from_numpy(d[k]).to(dev) for k in ("T0", "colp"))
    glob = d["glob"]                                                    # four host scalars
    Tout = torch.empty((B, N_SAVE, 32), device=dev, dtype=torch.float32)
    na, nr, st = (torch.zeros(B, dtype=torch.int32, device=dev) for _ in range(3))
    tstream = torch.cuda.Stream(device=dev)         # a real (non-NULL) stream: NULL would mean "the handle's own stream"
    torch.cuda.set_stream(tstream)
    stream = tstream.cuda_stream

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident throughput (`value`) ----------------
    sampler = ClockSampler(local_rank)
    sampler.start()                                  # before the warm-up: nvidia-smi needs a few hundred ms until its first sample
    for _ in range(warmup):
        h.solve_dev(T0, colp, glob, Tout, na, nr, st, stream=stream)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kernel_ms = []
    barrier()
    tw0 = time.time()
    e0.record()
    for _ in range(args.steps):
        h.solve_dev(T0, colp, glob, Tout, na, nr, st, stream=stream)
    e1.record()
    barrier()
    tw1 = time.time()
    total_ms = e0.elapsed_time(e1)
    window = "timed region"
    if sampler.count(tw0, tw1 + 0.05) < 2:
        # the timed region (steps x 10 ms) is shorter than nvidia-smi's sampling period: keep the SAME load running right behind it until
        # at least three samples have arrived (1.5 s at most) and report those
        tw0 = time.time()
        while time.time() - tw0 < 1.5 and sampler.count(tw0 + 0.1, time.time()) < 3:
            for _ in range(4):
                h.solve_dev(T0, colp, glob, Tout, na, nr, st, stream=stream)
            torch.cuda.synchronize()
        tw0, tw1 = tw0 + 0.1, time.time()
        window = f"the same solve repeated for {tw1 - tw0 + 0.1:.1f} s right behind the timed region ({total_ms:.0f} ms: shorter than nvidia-smi's sampling period)"
    clocks = sampler.stop(tw0, tw1, window)
    cnt = h.counters()                                      # counters + CUDA-event time of the LAST step's kernels
    kernel_ms = cnt["last_kernel_ms"]
    steps_per_pass = cnt["steps_accepted"] + cnt["steps_rejected"]
    assert int(st.abs().sum()) == 0, "solver status != 0 in the bench workload"
    t = torch.tensor([total_ms], device=dev, dtype=torch.float64)
    s = torch.tensor([float(steps_per_pass)], device=dev, dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(s, op=dist.ReduceOp.SUM)
    total_ms_max, steps_all = float(t.item()), float(s.item())
    value = steps_all * args.steps / (total_ms_max * 1e-3)

    # ---------------- end to end through the host-buffer C ABI (`e2e`) ----------------
    hT0 = torch.from_numpy(d["T0"]).pin_memory()
    hcolp = torch.from_numpy(d["colp"]).pin_memory()
    hglob = d["glob"]
    hout = torch.empty((B, N_SAVE, 32), dtype=torch.float32).pin_memory()
    h.solve(hT0.numpy(), hcolp.numpy(), hglob, out=hout.numpy())                # warm (allocates staging)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        r = h.solve(hT0.numpy(), hcolp.numpy(), hglob, out=hout.numpy())
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    e2e_steps = int((r["naccept"] + r["nreject"]).sum())
    te = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
    se = torch.tensor([float(e2e_steps)], device=dev, dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
        dist.all_reduce(se, op=dist.ReduceOp.SUM)
    e2e_value = float(se.item()) * args.steps / float(te.item())
    h2d = hT0.numel() * 4 + hcolp.numel() * 4
    d2h = hout.numel() * 4 + 3 * B * 4

    # ---------------- roofline of the dominant kernel (solve_tc_kernel) ----------------
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    bf16_peak = peaks.get("bf16_tflops", 1590.0)
    use_tc = os.environ.get("NFC_TC", "1") != "0"
    roofline = fp32_peak = None
    if rank == 0:
        fp32_peak = nfc_b200.measure_fp32_peak(local_rank, 50.0)                   # TFLOP/s, measured on this GPU now
        achieved = FLOPS_PER_COLUMN_STEP * steps_per_pass / (kernel_ms * 1e-3) / 1e12   # algorithmic (FP32-equivalent) TFLOP/s
        alg_bytes = B * (128 + 12 + N_SAVE * 128 + 12)
        hbm = {"achieved_gbs": alg_bytes / (kernel_ms * 1e-3) / 1e9, "peak_gbs": peaks.get("hbm_gbs", 6650.0),
               "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback", "algorithmic_bytes": alg_bytes}
        # DRAM traffic of the dominant kernel: from the ncu capture of the SAME kernel build (tools/ncu_traffic.py writes the file), else null
        traffic, traffic_source = None, "no ncu capture of this build under profiles/"
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "r02_solve_tc_traffic.json")))
            traffic, traffic_source = tj["dram_bytes_per_launch"], tj["source"]
        except (OSError, KeyError, ValueError):
            pass
        if use_tc:
            # tensor-core path: every fp32 product is three fp16 (default, 2-way split) or six bf16 (3-way split) tcgen05 MMAs, so the
            # ceiling for FP32-faithful work is the measured dense 16-bit peak / 3 (or / 6); burst figure: the kernel is timed alone
            nsplit = 6.0 if os.environ.get("NFC_TC_SPLIT", "fp16") == "bf16" else 3.0      # MMAs per fp32 product: BF16x3 -> 6, FP16x2 (default) -> 3
            roofline = {"bound": "tensor", "achieved": achieved, "peak": bf16_peak / nsplit, "unit": "TFLOP/s", "frac": achieved / (bf16_peak / nsplit),
                        "traffic": traffic, "traffic_source": traffic_source,
                        "peak_source": ("MEASURED_PEAKS.json bf16_tflops (burst)" if peaks else "fallback 1590") +
                                       (" / 6 MMAs per fp32 product (BF16x3 split)" if nsplit == 6.0 else " / 3 MMAs per fp32 product (FP16x2 split; fp16 and bf16 MMAs run at the same rate)"),
                        "tensor_tflops_16bit": nsplit * achieved, "fp32_simt_peak_measured": fp32_peak, "frac_of_fp32_simt_peak": achieved / fp32_peak,
                        "flops_per_column_step": FLOPS_PER_COLUMN_STEP, "kernel": "solve_tc_kernel (+ rhs_init_kernel, scheduler)", "kernel_ms": kernel_ms,
                        "hbm": hbm}
        else:
            roofline = {"bound": "fp32_fma", "achieved": achieved, "peak": fp32_peak, "unit": "TFLOP/s", "frac": achieved / fp32_peak, "traffic": traffic,
                        "peak_source": "measured in this run: pure-FFMA kernel in libnfc_b200 (nominal 148x128x2x1.965GHz = 74.4)",
                        "flops_per_column_step": FLOPS_PER_COLUMN_STEP, "kernel": "solve_kernel (+ rhs_init_kernel)", "kernel_ms": kernel_ms, "hbm": hbm}

    # ---------------- the schedule does not flatter the number: a handle without step history, and the first call of a fresh handle ----------------
    cold = None
    if rank == 0:
        try:
            os.environ["NFC_SCHED"] = "0"
            h2 = nfc_b200.Handle(cw, layers, 1, sv, device=local_rank)
            del os.environ["NFC_SCHED"]
            h2.set_weights(theta)
            torch.cuda.synchronize()
            e0.record()
            h2.solve_dev(T0, colp, glob, Tout, na, nr, st, stream=stream)          # first call of a fresh handle: scratch allocations inside
            e1.record()
            torch.cuda.synchronize()
            first_ms = e0.elapsed_time(e1)
            for _ in range(2):
                h2.solve_dev(T0, colp, glob, Tout, na, nr, st, stream=stream)
            torch.cuda.synchronize()
            e0.record()
            h2.solve_dev(T0, colp, glob, Tout, na, nr, st, stream=stream)
            e1.record()
            torch.cuda.synchronize()
            ns_ms = e0.elapsed_time(e1)
            cold = {"first_call_of_a_fresh_handle": {"ms": first_ms, "value": steps_per_pass / (first_ms * 1e-3)},
                    "no_longest_first_schedule": {"ms": ns_ms, "value": steps_per_pass / (ns_ms * 1e-3), "how": "NFC_SCHED=0: columns served in input order"},
                    "unit": UNIT}
            h2.close()
        except Exception as e:
            cold = {"error": str(e)}

    # ---------------- BASELINE config 5 (architecture sweep): the other Chains' forward solve on the same 65 536 columns ----------------
    other_chains = None
    if rank == 0 and world == 1:
        other_chains = {}
        for arch in ("conv-2", "conv-4", "dense-deeper", "dense-wider"):
            try:
                NNa = nfc_b200.named_chain(arch, seed=0)
                for p_ in NNa.params():
                    p_ *= 1e-5
                tha, _ = nfc_b200.destructure(NNa)
                ha = nfc_b200.Handle(NNa.conv_width, NNa.dense_spec, 1, sv, device=local_rank)
                ha.set_weights(tha)
                for _ in range(2):
                    ha.solve_dev(T0, colp, glob, Tout, na, nr, st, stream=stream)
                torch.cuda.synchronize()
                e0.record()
                ha.solve_dev(T0, colp, glob, Tout, na, nr, st, stream=stream)
                e1.record()
                torch.cuda.synchronize()
                ca = ha.counters()
                ms_a = e0.elapsed_time(e1)
                nsteps_a = int(ca["steps_accepted"] + ca["steps_rejected"])
                other_chains[arch] = {"ms": ms_a, "value": nsteps_a / (ms_a * 1e-3), "unit": UNIT,
                                      "kernel_path": "tcgen05" if (use_tc and arch != "dense-wider") else "simt", "bad_status": int((st != 0).sum().item())}
                ha.close()
            except Exception as e:
                other_chains[arch] = {"error": str(e)}

    # ---------------- second BASELINE metric: adjoint gradient, seconds per epoch (rank 0, one GPU) ----------------
    # epoch = one loss + gradient over the batch (src/training.jl:45-48,70): (a) BASELINE config 3, the 9 training simulations (one column per
    # CTA, FP32); (b) the config-2 batch (65536 columns): the tensor-core backward kernel (csrc/nfc_adjoint_tc.cuh).
    def adjoint_roofline(ca, ms):
        tf_ = (ca["adj_steps_accepted"] + ca["adj_steps_rejected"]) * FLOPS_PER_ADJOINT_COLUMN_STEP / (ms * 1e-3) / 1e12
        if use_tc and os.environ.get("NFC_ADJ_TC", "1") != "0":
            return {"bound": "tensor", "achieved": tf_, "peak": bf16_peak / 3.0, "unit": "TFLOP/s", "frac": tf_ / (bf16_peak / 3.0),
                    "flops_per_adjoint_column_step": FLOPS_PER_ADJOINT_COLUMN_STEP, "frac_of_fp32_simt_peak": tf_ / fp32_peak if fp32_peak else None,
                    "kernel": "adjoint_tc_kernel (tcgen05: forward recompute + transposed chain FP16x2, dW contraction BF16) + the tensor-core RECORD "
                              "forward pass, both inside device_ms",
                    "peak_source": "MEASURED_PEAKS.json bf16_tflops (burst) / 3 MMAs per fp32 product" if peaks else "fallback 1590 / 3"}
        return {"bound": "fp32_fma", "achieved": tf_, "peak": fp32_peak, "unit": "TFLOP/s", "frac": tf_ / fp32_peak if fp32_peak else None,
                "flops_per_adjoint_column_step": FLOPS_PER_ADJOINT_COLUMN_STEP, "kernel": "adjoint_kernel (FP32 SIMT) + RECORD forward pass"}

    adjoint = None
    if rank == 0:
        try:
            syn = nfc_b200.synthetic
            adjoint = {}
            for tag, dd in (("config3_9sims", syn.make_columns("training9")), ("ocean_65536", syn.make_columns("ocean", B=65536, seed=2))):
                nb = dd["T0"].shape[0]
                Tt = torch.from_numpy(np.repeat(dd["T0"][:, None, :], N_SAVE, axis=1) * np.float32(0.98)).to(dev)
                a0, ac = torch.from_numpy(dd["T0"]).to(dev), torch.from_numpy(dd["colp"]).to(dev)
                g = torch.empty(h.n_params, device=dev)
                ls = torch.zeros(2, dtype=torch.float64, device=dev)
                sta = torch.zeros(nb, dtype=torch.int32, device=dev)
                for it in range(3):
                    torch.cuda.synchronize()
                    t0 = time.perf_counter()
                    h.loss_grad_dev(a0, ac, dd["glob"], Tt, ls, g, sta, n_total_columns=nb, stream=stream)
                    torch.cuda.synchronize()
                    dt_ = time.perf_counter() - t0
                ca = h.counters()
                adjoint[tag] = {"sec_per_epoch": dt_, "columns": nb, "fwd_column_steps": ca["steps_accepted"] + ca["steps_rejected"],
                                "adjoint_column_steps": ca["adj_steps_accepted"] + ca["adj_steps_rejected"], "device_ms": ca["last_adjoint_ms"],
                                "record_pass_ms": ca["last_kernel_ms"], "bad_status": int((sta != 0).sum()),
                                "adjoint_column_steps_per_sec": (ca["adj_steps_accepted"] + ca["adj_steps_rejected"]) / max(ca["last_adjoint_ms"] * 1e-3, 1e-9)}
                if nb >= 8192:
                    adjoint[tag]["roofline"] = adjoint_roofline(ca, ca["last_adjoint_ms"])
                del Tt
                if tag == "config3_9sims" and not args.no_cpu_baseline:
                    # the CPU twin of the same continuous adjoint (oracle/nde_oracle.c, OpenMP over the 9 simulations) on the box's cores
                    from oracle import c_oracle as CO
                    Tth = np.repeat(dd["T0"][:, None, :], N_SAVE, axis=1) * np.float32(0.98)
                    cores = min(os.cpu_count() or 1, nb)
                    tc_ = []
                    for it in range(2):
                        t0 = time.perf_counter()
                        rc_ = CO.loss_grad_continuous(dd["T0"], theta, cw, layers, dd["colp"], dd["glob"], sv, Tth, nthreads=cores)
                        tc_.append(time.perf_counter() - t0)
                    adjoint[tag]["cpu_port"] = {"sec_per_epoch": min(tc_), "cores": cores, "kind": "port",
                                                "adjoint_column_steps": int((rc_["adj_naccept"] + rc_["adj_nreject"]).sum())}
        except Exception as e:                                     # the forward metric must not die with the extra one
            adjoint = {"error": str(e)}

    # ---------------- BASELINE config 4: 1,048,576 columns (0.25 degree ocean), 129 saves, sharded over the N ranks (STRONG scaling) ----------------
    # forward solve (no collective), then loss + gradient with the all-reduce of [grad; loss sums] inside the timed region: the collective runs
    # inside nfc_loss_grad_dev on the handle's own NCCL communicator (torch.distributed only ships the 128-byte rendezvous id).
    config4 = None
    if use_tc and not args.no_large_batch:
        try:
            del Tout, hout
            lo, hi = nfc_b200.distributed.shard_range(CONFIG4_COLUMNS, rank, world)
            nb = hi - lo
            # every rank generates the same 1M-column set in pieces and keeps its shard (identical columns whatever N is)
            dl = nfc_b200.synthetic.make_columns("ocean", B=CONFIG4_COLUMNS, seed=2, K_CA=2.0)
            lT0, lcolp = (torch.from_numpy(np.ascontiguousarray(dl[k][lo:hi])).to(dev) for k in ("T0", "colp"))
            lglob = dl["glob"]
            del dl
            lout = torch.empty((nb, N_SAVE, 32), device=dev, dtype=torch.float32)
            lst = torch.zeros(nb, dtype=torch.int32, device=dev)
            for _ in range(2):
                h.solve_dev(lT0, lcolp, lglob, lout, None, None, lst, stream=stream)
            barrier()
            e0.record()
            h.solve_dev(lT0, lcolp, lglob, lout, None, None, lst, stream=stream)
            e1.record()
            barrier()
            cl = h.counters()
            fms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
            fst = torch.tensor([float(cl["steps_accepted"] + cl["steps_rejected"]), float((lst != 0).sum())], device=dev, dtype=torch.float64)
            if dist is not None:
                dist.all_reduce(fms, op=dist.ReduceOp.MAX)
                dist.all_reduce(fst, op=dist.ReduceOp.SUM)
            fwd_ms, fwd_steps = float(fms.item()), float(fst[0].item())
            ftf = fwd_steps * FLOPS_PER_COLUMN_STEP / (fwd_ms * 1e-3) / 1e12
            config4 = {"workload": f"config4: {CONFIG4_COLUMNS} columns sharded over {world} GPU(s) ({nb} on rank 0), 129 saves, same Chain / tolerances",
                       "scaling": "strong", "columns_per_rank": nb,
                       "forward": {"ms": fwd_ms, "column_steps": fwd_steps, "value": fwd_steps / (fwd_ms * 1e-3), "unit": UNIT, "bad_status": int(fst[1].item()),
                                   "roofline": {"bound": "tensor", "achieved": ftf, "peak": world * bf16_peak / 3.0, "unit": "TFLOP/s", "frac": ftf / (world * bf16_peak / 3.0),
                                                "kernel": "solve_tc2_kernel (two 128-column tiles in flight per CTA; picked from 98304 columns per call on)"}}}
            # loss + gradient: truth = the forward solution perturbed in closed form on the device (no 17 GB host read)
            lTt = lout.mul_(0.99).add_(0.01)
            del lout
            if dist is not None:
                nfc_b200.distributed.attach_communicator(h)
            g = torch.empty(h.n_params, device=dev)
            ls = torch.zeros(2, dtype=torch.float64, device=dev)
            times = []
            # adjoint tape: 344 KB per column as a dense tape (360 GB for this batch on one GPU) -> compact layout (tape_mode 0 picks it when the
            # dense tape exceeds the budget): ~ 87 KB per column, three chunks inside the default 48 GiB budget on one GPU (one chunk with
            # tape_gib = 100 is 0.5 % faster: profiles/r02_tape_layouts.txt)
            tape_gib = 48
            free_before = torch.cuda.mem_get_info(dev)[0]
            for it in range(3):
                barrier()
                e0.record()
                h.loss_grad_dev(lT0, lcolp, lglob, lTt, ls, g, lst, n_total_columns=CONFIG4_COLUMNS, stream=stream)
                e1.record()
                barrier()
                times.append(e0.elapsed_time(e1))
            ca = h.counters()
            ams = torch.tensor([times[-1], ca["last_allreduce_ms"]], device=dev, dtype=torch.float64)
            ast = torch.tensor([float(ca["adj_steps_accepted"] + ca["adj_steps_rejected"]), float(ca["steps_accepted"] + ca["steps_rejected"]),
                                float((lst != 0).sum())], device=dev, dtype=torch.float64)
            if dist is not None:
                dist.all_reduce(ams, op=dist.ReduceOp.MAX)
                dist.all_reduce(ast, op=dist.ReduceOp.SUM)
            lsh = ls.cpu()
            atf = float(ast[0].item()) * FLOPS_PER_ADJOINT_COLUMN_STEP / (float(ams[0].item()) * 1e-3) / 1e12
            config4["loss_grad"] = {"sec_per_epoch": float(ams[0].item()) * 1e-3, "first_call_sec": times[0] * 1e-3,
                                    "tape": {"layout": "compact (rows per column from the previous call's step counts; first call: counting pass)" if nb * 512 * 672 > tape_gib * 2**30 else "dense",
                                             "budget_gib": tape_gib, "scratch_gib_rank0": (free_before - torch.cuda.mem_get_info(dev)[0]) / 2**30,
                                             "dense_tape_would_need_gib": nb * (512 * 672 + N_SAVE * 128) / 2**30},
                                    "allreduce_us_max_over_ranks": float(ams[1].item()) * 1e3 if dist is not None else 0.0,
                                    "collective": ("one grouped ncclAllReduce of 24735 floats + 2 doubles inside nfc_loss_grad_dev (C-ABI communicator), inside the timed region"
                                                   if dist is not None else "none (one GPU)"),
                                    "adjoint_column_steps": float(ast[0].item()), "fwd_column_steps": float(ast[1].item()), "bad_status": int(ast[2].item()),
                                    "adjoint_column_steps_per_sec": float(ast[0].item()) / (float(ams[0].item()) * 1e-3),
                                    "loss": float(lsh[0] / lsh[1]) if dist is not None else float(lsh[0] / lsh[1]),
                                    "grad_norm": float(g.double().norm().item()),
                                    "roofline": {"bound": "tensor", "achieved": atf, "peak": world * bf16_peak / 3.0, "unit": "TFLOP/s", "frac": atf / (world * bf16_peak / 3.0),
                                                 "flops_per_adjoint_column_step": FLOPS_PER_ADJOINT_COLUMN_STEP,
                                                 "note": "whole loss+gradient call (RECORD forward pass, backward kernel, tape chunking, all-reduce) against the tensor ceiling of all ranks"}}
            del lTt, lT0, lcolp, lst
        except Exception as e:
            config4 = {"error": f"{type(e).__name__}: {e}"}

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return 0

    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        sample = 2048
        v, ms, nst = cpu_reference_run(1, 1, sample, cores)
        v1, ms1, nst1 = cpu_reference_run(1, 0, 256, 1)          # the reference's actual mode: one thread (train_free_convection_nde.jl:23)
        cpu_baseline = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                        "sample": f"{sample} of the {B} columns, {nst} column-steps, {ms:.0f} ms, C port of the reference path (OpenMP over columns)",
                        "single_thread": {"value": v1, "unit": UNIT, "cores": 1, "sample": f"256 of the {B} columns, {nst1} column-steps, {ms1:.0f} ms"}}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": warmup,
            "ms_per_step": total_ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": ("f32 (fp16x2-split tcgen05 MMAs, fp32 accumulate)" if os.environ.get("NFC_TC_SPLIT", "fp16") != "bf16" else "f32 (bf16x3-split tcgen05 MMAs, fp32 accumulate)") if use_tc else "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "kernel_path": "tcgen05" if use_tc else "simt", "tiles_in_flight_per_cta": (2 if (os.environ.get("NFC_TC_TILES", "0") == "2" or (os.environ.get("NFC_TC_TILES", "0") == "0" and B >= int(os.environ.get("NFC_TC2_MIN", "98304")))) else 1) if use_tc else None, "columns_per_gpu": B, "column_steps_per_pass_rank0": steps_per_pass,
                       "l2": "outputs (1.08 GB/step) exceed the 126 MB L2; no explicit flush", "parallelism": f"columns sharded x{world}, no collective"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": int(cnt["kernel_launches"]) * args.steps, "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu_baseline, "adjoint": adjoint,
            "config4": config4, "cold_call": cold, "other_chains": other_chains, "cpu_affinity": affinity}
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
