/* nfc.h -- C ABI of libnfc_b200.so: the B200 (sm_100a) hot path of NeuralFreeConvection.jl.
 *
 * The reference has no FFI; its seam is the SciML protocol (an out-of-place ODEProblem whose
 * f(u,p,t) closes over the operators, solved with u0=/p= overrides).  Each entry point below
 * replaces the numerical work behind one reference call site (file:line into the reference):
 *
 *   nfc_create / nfc_destroy   FreeConvectionNDE(NN, ds; iterations)        src/free_convection_nde.jl:1-47
 *                              ConvectiveAdjustmentNDE(NN, ds; iterations)  src/convective_adjustment_nde.jl:1-57
 *   nfc_set_weights            Flux.destructure(NN) -> theta                src/solve.jl:2, src/training.jl:43
 *   nfc_rhs                    the closure dT/dt(T, p, t)                   src/free_convection_nde.jl:29-38,
 *                                                                           src/convective_adjustment_nde.jl:33-48
 *   nfc_solve                  solve_nde(nde, NN, T0, alg, nde_params)      src/solve.jl:1-6
 *   nfc_diagnose_flux          flux diagnosis loop of solve_nde(ds, ...)    src/solve.jl:32-48
 *   nfc_loss_grad              nde_loss + Zygote gradient                   src/training.jl:45-48,70
 *
 * Conventions
 *   - plain C, host pointers unless the name ends in _dev; the caller owns every buffer for the
 *     duration of the call only; the handle owns device memory, streams, scratch.
 *   - every call returns 0 on success, negative on usage / CUDA error (text via nfc_last_error).
 *     Per-column numerical trouble is reported in status[] (0 ok, 1 maxiters, 2 NaN/Inf, 3 dt underflow,
 *     4 adjoint tape overflow, 5 operand range of the FP16x2 tensor-core split left: |T_scaled| > 128 -- the host-buffer
 *     calls re-solve such columns without that bound, see `range_fallback`), never as a call failure -- the reference
 *     treats solver retcodes as warnings.
 *   - calls on one handle are serialised by the caller; host-pointer calls are synchronous.
 *   - there is NO CPU fallback: without a CUDA device nfc_create fails.
 *
 * Layouts (Julia column-major == C row-major with reversed dims)
 *   T, dTdt, T0   [B][Nz]           Julia Nz x B;      index 0 = deepest cell, already T_scaling'ed
 *   colp          [B][3]            (bottom_flux, top_flux, K_CA) per column, wT_scaling'ed fluxes
 *                                   (FreeConvectionNDEParameters / ConvectiveAdjustmentNDEParameters,
 *                                    src/free_convection_nde.jl:49-62, src/convective_adjustment_nde.jl:59-72)
 *   glob          [4]               (sigma_T, sigma_wT, H, tau)
 *   theta         [n_params]        Flux.destructure order: [conv weight(k), conv bias(1)], then per Dense:
 *                                   weight (out x in, column-major), bias
 *   Tout, Ttrue   [B][n_save][Nz]   Julia Nz x n_save x B
 *   wT            [B][n_save][Nz+1]
 */
#ifndef NFC_H
#define NFC_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define NFC_MAX_LAYERS 8
#define NFC_ACT_IDENTITY 0
#define NFC_ACT_RELU 1
#define NFC_NDE_FREE_CONVECTION 0      /* K_CA column of colp is ignored (treated as 0) */
#define NFC_NDE_CONVECTIVE_ADJUSTMENT 1
#define NFC_ALG_TSIT5 0
#define NFC_ALG_RKC2 1   /* stabilised explicit Runge-Kutta-Chebyshev, in the role of ROCK4 (train_free_convection_nde.sh:3); nfc_solve only */

typedef struct nfc_handle nfc_handle;

typedef struct nfc_config {
    int32_t struct_size;                          /* = sizeof(nfc_config), ABI guard */
    int32_t Nz;                                   /* grid points (--grid-points, train_free_convection_nde.jl:29,90): 4 .. 32.  32 = every kernel.
                                                   * Nz < 32: the same five Chains built on Nz (Dense(Nz, 4 Nz) ... Dense(4 Nz, Nz - 1)) through the
                                                   * host-buffer calls nfc_rhs / nfc_solve / nfc_diagnose_flux / nfc_loss_grad (Tsit5, FP32 kernels: the
                                                   * Chain is embedded zero-padded in its 32-level version); all shapes and theta are the Nz-level ones */
    int32_t conv_width;                           /* 0 = none, else Conv((k,1),1=>1,relu) front-end (2 or 4) */
    int32_t n_layers;                             /* Dense layers */
    int32_t layer_dims[NFC_MAX_LAYERS + 1];       /* in_0, out_0, out_1, ... */
    int32_t activations[NFC_MAX_LAYERS];          /* NFC_ACT_* per Dense layer */
    int32_t nde_type;                             /* NFC_NDE_* */
    int32_t alg;                                  /* NFC_ALG_TSIT5 | NFC_ALG_RKC2 */
    float reltol;                                 /* src/solve.jl:4 -> 1e-4 */
    float abstol;                                 /* DiffEqBase default 1e-6 */
    float fixed_dt;                               /* > 0: adaptive=false, dt=fixed_dt (test / fixed-step use) */
    int32_t max_iters;                            /* DiffEqBase default 100000 */
    int32_t n_save;                               /* length(saveat) */
    const float* saveat;                          /* [n_save] ascending, saveat[0] = t0 = 0, saveat[n_save-1] = tf */
    int32_t device;                               /* CUDA device ordinal */
    int32_t adjoint_max_steps;                    /* accepted forward steps kept per column for nfc_loss_grad (0 = default 512; overflow -> status 4) */
    float adjoint_reltol;                         /* 0 = reltol */
    float adjoint_abstol;                         /* 0 = abstol */
    int32_t reserved[8];
} nfc_config;

typedef struct nfc_counters {
    int64_t columns;            /* columns processed by the last solve / loss_grad call */
    int64_t steps_accepted;     /* forward Tsit5 steps accepted, all columns */
    int64_t steps_rejected;     /* forward Tsit5 steps rejected */
    int64_t adj_steps_accepted; /* backward (adjoint) steps accepted */
    int64_t adj_steps_rejected;
    int64_t kernel_launches;    /* kernels launched by the last call */
    float   last_kernel_ms;     /* device time of the last call's main kernel(s), CUDA events */
    float   last_adjoint_ms;
    int64_t rhs_evals;          /* forward RHS evaluations of the last solve (Tsit5: 6 per attempted step + 2 per column) */
    float   last_allreduce_ms;  /* device time of the [grad; loss] all-reduce of the last nfc_loss_grad (0 without a communicator) */
    int32_t comm_world;         /* ranks of the attached communicator (1: none) */
} nfc_counters;

int32_t nfc_create(const nfc_config* cfg, nfc_handle** out);
int32_t nfc_destroy(nfc_handle* h);
const char* nfc_last_error(const nfc_handle* h);   /* h may be NULL: error of the last failed nfc_create */
int64_t nfc_n_params(const nfc_handle* h);
int32_t nfc_set_weights(nfc_handle* h, const float* theta, int64_t n);
int32_t nfc_set_saveat(nfc_handle* h, const float* saveat, int32_t n_save);
int32_t nfc_get_counters(const nfc_handle* h, nfc_counters* out);

/* host-buffer calls (the drop-in boundary).  A column whose status is not 0 stops saving where it failed: nfc_solve returns NaN in the
 * remaining rows of its trajectory (the *_dev calls leave those rows of the caller's buffer untouched). */
int32_t nfc_rhs(nfc_handle* h, const float* T, const float* colp, const float* glob, float* dTdt, int64_t B);
int32_t nfc_solve(nfc_handle* h, const float* T0, const float* colp, const float* glob, float* Tout,
                  int32_t* naccept, int32_t* nreject, int32_t* status, int64_t B);
int32_t nfc_diagnose_flux(nfc_handle* h, const float* Tsol, const float* colp, float* wT, int64_t B);
int32_t nfc_loss_grad(nfc_handle* h, const float* T0, const float* colp, const float* glob, const float* Ttrue,
                      double* loss, float* grad_theta, int32_t* status, int64_t B);

/* device-buffer calls: same semantics, array pointers are device memory on cfg.device (glob stays a HOST pointer:
 * four scalars), work is enqueued on `stream` (a cudaStream_t passed as void*; NULL = the handle's own stream, which is
 * non-blocking and therefore NOT ordered with the legacy default stream: a host whose buffers are produced on the default
 * stream passes cudaStreamLegacy, (void*)0x1, or cudaStreamPerThread, (void*)0x2) and NOT synchronised. */
int32_t nfc_rhs_dev(nfc_handle* h, const float* T, const float* colp, const float* glob, float* dTdt, int64_t B, void* stream);
int32_t nfc_solve_dev(nfc_handle* h, const float* T0, const float* colp, const float* glob, float* Tout,
                      int32_t* naccept, int32_t* nreject, int32_t* status, int64_t B, void* stream);
/* loss_sum_dev: double[2] = (sum of squared residuals, number of residuals) of this rank's columns;
 * grad_dev: float[n_params] = d(sum of squared residuals / n_total)/dtheta with n_total = n_total_columns*n_save*Nz,
 * so that a SUM all-reduce over ranks yields the global-batch gradient (src/training.jl:47 mean over all sims).
 * With a communicator attached (nfc_comm_init) the all-reduce happens inside the call: loss_sum_dev and grad_dev then hold the GLOBAL
 * sums / gradient on every rank and n_total_columns is ignored. */
int32_t nfc_loss_grad_dev(nfc_handle* h, const float* T0, const float* colp, const float* glob, const float* Ttrue,
                          double* loss_sum_dev, float* grad_dev, int32_t* status, int64_t B, int64_t n_total_columns,
                          void* stream);
int32_t nfc_sync(nfc_handle* h);

/* Multi-GPU (the reference has none): one handle per GPU (one process per GPU, or several handles in one process), columns
 * sharded over the ranks by the host, NO data-path collective in the forward solve.  A handle with a communicator performs the ONE
 * collective of a training step itself: nfc_loss_grad / nfc_loss_grad_dev then take this rank's columns and return the loss and the
 * gradient of the GLOBAL batch on every rank (src/training.jl:45-48,70 over all ranks' simulations) -- one grouped ncclAllReduce of
 * [dL/dtheta (n_params floats); sum of squared residuals, residual count (2 doubles)] over NVLink, enqueued behind the backward kernel.
 * NCCL is loaded at run time (dlopen "libnccl.so.2": the library the host process already uses, else the system one).
 *   nfc_comm_unique_id  rank 0 creates the 128-byte rendezvous id; the HOST ships it to the other ranks (MPI, a file, a socket)
 *   nfc_comm_init       collective over all ranks: attaches rank `rank` of `world` to the handle (world = 1 is allowed)
 *   nfc_comm_destroy    detaches (nfc_destroy does it too) */
#define NFC_UNIQUE_ID_BYTES 128
int32_t nfc_comm_unique_id(void* id128);
int32_t nfc_comm_init(nfc_handle* h, const void* id128, int32_t rank, int32_t world);
int32_t nfc_comm_destroy(nfc_handle* h);

/* SURVEY row f1 -- the embedded path oceananigans_convective_adjustment[_with_neural_network] (src/oceananigans_nn.jl:97-282):
 * n_steps fixed steps of dt: explicit neural-network flux step in physical units (forward Euler stands in for Oceananigans' own
 * time stepper), then the implicit convective adjustment of convective_adjustment! (src/oceananigans_nn.jl:3-30) as one tridiagonal
 * solve per column.  T0 [B][Nz] and Tout [B][n_steps/save_every + 1][Nz] are PHYSICAL temperatures; heat_flux [B] is the surface
 * temperature flux; scal = (mu_T, sigma_T, mu_wT, sigma_wT) of the two ZeroMeanUnitVarianceScalings; K is the diffusivity as the
 * reference passes it to convective_adjustment! (already multiplied by sigma_wT/sigma_T*stop_time/Lz in the NN variant, :195).
 * Set the weights to zero for the pure convective-adjustment baseline.  Host pointers; synchronous. */
int32_t nfc_embedded_solve(nfc_handle* h, const float* T0, const float* heat_flux, const float* scal, float Lz, float dt, float K,
                           float bottom_gradient, int32_t n_steps, int32_t save_every, float* Tout, int64_t B);
/* second output of the embedded path, diagnose_wT_NN (src/oceananigans_nn.jl:200-218): for saved PHYSICAL profiles Tsol [B][n_out][Nz]
 * (what nfc_embedded_solve returned) wT [B][n_out][Nz+1] = [0; wT_scaling^-1(NN(T_scaling(T))); heat_flux] - kappa dT/dz on the interior
 * faces, kappa = K where dT/dz < 0 (K as passed to nfc_embedded_solve). */
int32_t nfc_embedded_diagnose_flux(nfc_handle* h, const float* Tsol, const float* heat_flux, const float* scal, float Lz, float K,
                                   int32_t n_out, float* wT, int64_t B);
/* SURVEY row f3 -- compute_nde_solution_history (src/testing.jl:42-79): every saved epoch's Chain x every simulation, which the
 * reference runs as epochs x simulations serial solve_nde calls.  Here it is one launch over an ensemble: member e (one epoch's
 * Flux.destructure vector theta_members[e][n_params], same Chain architecture as the handle) integrates the S columns
 * T0[e*S .. (e+1)*S) with its own weights; Tout [E*S][n_save][Nz].  The handle's own weights (nfc_set_weights) are neither needed
 * nor touched.  Columns follow exactly the arithmetic of nfc_solve with the FP32 kernels (NFC_TC=0).  Host pointers; synchronous. */
int32_t nfc_solve_ensemble(nfc_handle* h, const float* theta_members, int64_t n_members, const float* T0, const float* colp,
                           const float* glob, float* Tout, int32_t* naccept, int32_t* nreject, int32_t* status,
                           int64_t columns_per_member);
/* diagnostics: attempted backward (adjoint) steps per column of the last nfc_loss_grad chunk (the analogue of naccept + nreject) */
int32_t nfc_get_adjoint_steps(nfc_handle* h, int32_t* out, int64_t n);

/* Kernel selection and scratch budgets (defaults are chosen by batch size; the same names in upper case with the prefix NFC_ are read
 * from the environment at nfc_create / call time and win over the defaults):
 *   tc               0: FP32 SIMT kernels; 1: tensor-core kernels (the default where they exist: every call for the default Chain; nfc_rhs and
 *                    the Tsit5 forward solve for conv-2 / conv-4 / deeper; none for the wider Chain)
 *   tc_tiles         0: pick by batch size; 1 / 2: one / two 128-column tiles in flight per CTA in the forward solve
 *   tc2_min_columns  batch size from which tc_tiles = 0 picks two tiles (default 98304)
 *   tc_record        0: forward (RECORD) pass of nfc_loss_grad on the FP32 SIMT kernel
 *   sched            0: columns served in input order; 1 (default): longest first from the previous call's step counts
 *   fwd_small_max    forward solves / RECORD passes of at most this many columns (default Chain) run one column per CTA (default 592 = four per
 *                    SM; 0 = never; forcing tc_tiles also bypasses it)
 *   adj_small_max    batches up to this many columns run the one-column-per-CTA backward kernel (default 2000, half of it once the
 *                    previous call's per-column step counts are known; 0 = never)
 *   adj_tc           0: batch backward pass on the FP32 SIMT kernel instead of the tensor-core kernel
 *   adj_heavy        how many of a chunk's longest columns (by the previous call's backward step counts) run on the one-column-per-CTA
 *                    kernel next to the tensor-core backward kernel (default 48, at most 1/16 of the chunk; 0 = none)
 *   e2e_pieces       nfc_solve: pieces whose device->host copy overlaps the next piece's solve (0 = auto: 2 from 16384, 3 from 49152 columns)
 *   tc_split         0 = FP16x2 operands (3 MMAs per fp32 product, needs |T_scaled| <= 128; default), 1 = BF16x3 (6 MMAs, fp32's range)
 *   range_fallback   1 (default): nfc_solve re-solves columns that ended with status 5 on the BF16x3 split, nfc_loss_grad (no communicator)
 *                    repeats a call that had such a column on the FP32 kernels; 0: status 5 is returned as it is (the *_dev calls always do)
 *   reset_history    (any value) forget the per-column step counts kept from the previous calls: schedules, side launch and compact-tape
 *                    shares are keyed by the batch size only -- call it when NEW columns arrive under an unchanged batch size (results never
 *                    depend on the history, only the time does: a stale one costs ~ 8 % on the next call)
 *   tape_gib         tape budget of nfc_loss_grad in GiB; larger batches are processed in chunks (0 = default 48)
 *   tape_mode        layout of the adjoint tape: 1 = dense (every column owns adjoint_max_steps rows of 672 B: 344 KB at the default 512),
 *                    2 = compact (a counting pass -- the forward solve on the RECORD pass's kernels, no outputs -- sizes every column's
 *                    share: accepted steps + 2 rows, ~ 75 KB per column at 110 steps; adjoint_max_steps does not apply; the call
 *                    synchronises its stream once to read the counts), 0 (default) = compact when the dense tape of the whole batch
 *                    exceeds tape_gib, else dense */
int32_t nfc_set_option(nfc_handle* h, const char* name, int64_t value);

/* diagnostics: the forward tape of one column of the last nfc_loss_grad chunk -- per accepted step 168 floats: u_n[32], then the Tsit5
 * dense output in power basis c1..c4 [4][32] (u(t_n + th h) = u_n + h th (c1 + th (c2 + th (c3 + th c4)))), t_n, h_n, 6 unused. */
int32_t nfc_get_tape(nfc_handle* h, int64_t column, float* rows, int32_t max_rows, int32_t* n_rows);

/* Roofline denominator for this path: sustained FP32 FMA throughput of `device`, measured with a
 * pure-FFMA kernel (16 independent chains per thread, full occupancy) over ~`ms_target` milliseconds. */
int32_t nfc_measure_fp32_peak(int32_t device, float ms_target, double* tflops);

#ifdef __cplusplus
}
#endif
#endif /* NFC_H */
