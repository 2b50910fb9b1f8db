// nfc_forward.cuh -- forward kernels: weight packing, batched RHS / initial step, persistent Tsit5 solve.
//
//   pack_weights_kernel   Flux.destructure theta -> lane-permuted packed weights          (src/solve.jl:2)
//   rhs_init_kernel       MODE 0: dT/dt for [B] columns                                   (src/free_convection_nde.jl:29-38,
//                                                                                          src/convective_adjustment_nde.jl:33-48)
//                         MODE 1: k1 = f(u0) and the Hairer initial dt per column         (OrdinaryDiffEq initdt [UPSTREAM])
//   solve_kernel          solve(nde, Tsit5(), reltol, u0=T0, p=[theta; nde_params]) with  (src/solve.jl:4)
//                         saveat served by dense output                                   (src/free_convection_nde.jl:40-41)
#pragma once
#include "nfc_device.cuh"

namespace nfc {

enum : int { ST_OK = 0, ST_MAXITERS = 1, ST_NAN = 2, ST_DT_UNDERFLOW = 3, ST_TAPE_OVERFLOW = 4, ST_RANGE = 5 };   // ST_RANGE: FP16x2 operand range left (|T_scaled| > 128)
enum : int { CTR_QUEUE = 0, CTR_ACCEPT = 1, CTR_REJECT = 2, CTR_ADJ_QUEUE = 3, CTR_ADJ_ACCEPT = 4, CTR_ADJ_REJECT = 5, CTR_ADJ_QUEUE_HEAVY = 6, CTR_HEAVY_STARTED = 7, CTR_COUNT = 8 };

struct SolveParams {
    const float* wpack;        // packed weights
    const float* T0;           // [B][32]
    const float* colp;         // [B][3]
    const float* k1_init;      // [B][32]  f(u0)
    const float* dt_init;      // [B]
    const float* saveat;       // [n_save]
    const int* order;          // [B] queue position -> column (longest-first schedule from the previous call) or null
    int* work;                 // [B] attempted steps per column of THIS call (feeds the next call's schedule) or null
    float* Tout;               // [B][n_save][32]
    int* naccept;              // [B] (may be null)
    int* nreject;
    int* status;
    unsigned long long* counters;
    long long B;
    int n_save;
    int max_iters;
    int use_kca;               // 0: FreeConvectionNDE (K_CA forced to 0)
    float pref, tf, reltol, abstol, fixed_dt;
    // record mode (loss / adjoint tape), see nfc_adjoint.cuh
    const float* Ttrue;        // [B][n_save][32]
    float* resid;              // [B][n_save][32]  u(t_i) - Ttrue
    float* tape;               // dense: [B][tape_cap][TAPE_STRIDE]; compact: rows [tape_off[c], tape_off[c + 1]) belong to column c
    int* tape_len;             // [B]
    double* loss_sum;          // [1]
    int tape_cap;
    int nz;                    // vertical levels in use (<= 32; 0 means 32): FP32 SIMT kernels only
    const long long* tape_off; // [B + 1] first tape row of every column (compact tape: a counting pass sized every column's share) or null
};

// Tape addressing shared by the RECORD and the backward kernels (SolveParams / AdjParams).
template <class PP> __device__ __forceinline__ long long tape_row0(const PP& P, long long col) { return P.tape_off ? P.tape_off[col] : col * P.tape_cap; }
template <class PP> __device__ __forceinline__ int tape_capacity(const PP& P, long long col) {
    return P.tape_off ? (int)(P.tape_off[col + 1] - P.tape_off[col]) : P.tape_cap;
}

template <class N>
__global__ void pack_weights_kernel(const float* __restrict__ theta, float* __restrict__ wpack) {
    static_assert(N::P_TOTAL % 4 == 0, "ensemble members are staged with 16-byte aligned bulk copies");
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= N::P_TOTAL) return;
    theta += (size_t)blockIdx.y * N::T_TOTAL;     // blockIdx.y = ensemble member (0 for the handle's own weights)
    wpack += (size_t)blockIdx.y * N::P_TOTAL;
    const int t = packed_to_theta<N>(p);
    wpack[p] = t >= 0 ? theta[t] : 0.f;
}

// ---------------------------------------------------------------------------------------------------
// One f evaluation for the warp's 8 columns. y in registers (lane == level) -> f in registers.
// ---------------------------------------------------------------------------------------------------
template <class N>
__device__ __forceinline__ void rhs_tile(const float* sW, float* ws, const float (&y)[CPW], float (&f)[CPW], const float* bot,
                                         const float* top, const float* kca, float pref, int lane, int nz = NZ) {
    nn_forward_tile<N>(sW, ws, y, lane);
    const float* sY = ws + N::S_Y;
#pragma unroll
    for (int c = 0; c < CPW; ++c) f[c] = flux_divergence(y[c], sY + c * YS, bot[c], top[c], kca[c], pref, lane, nz);
    // the next nn_forward_tile writes sX/sA first and reaches sY only after several __syncwarp()s
}

// f(T) (MODE 0) or k1 = f(u0) plus the Hairer initial dt (MODE 1) for the 8 columns [col0, col0 + 8) below colEnd.
template <class N, int MODE>
__device__ __forceinline__ void rhs_init_group(const float* sW, float* ws, long long col0, long long colEnd, const float* __restrict__ T,
                                               const float* __restrict__ colp, float* __restrict__ out, float* __restrict__ dt_out, float pref,
                                               int use_kca, float reltol, float abstol, float tf, float fixed_dt, int lane, int nz = NZ) {
    const float inz = 1.f / (float)nz;          // RMS norms run over the nz real levels (the padded ones hold zeros)
    float y[CPW], f0[CPW], bot[CPW], top[CPW], kca[CPW];
#pragma unroll
    for (int c = 0; c < CPW; ++c) {
        const long long col = col0 + c;
        const bool ok = col < colEnd;
        y[c] = ok ? T[col * NZ + lane] : 0.f;
        bot[c] = ok ? colp[col * 3 + 0] : 0.f;
        top[c] = ok ? colp[col * 3 + 1] : 0.f;
        kca[c] = (ok && use_kca) ? colp[col * 3 + 2] : 0.f;
    }
    rhs_tile<N>(sW, ws, y, f0, bot, top, kca, pref, lane, nz);
#pragma unroll
    for (int c = 0; c < CPW; ++c) {
        const long long col = col0 + c;
        if (col < colEnd) out[col * NZ + lane] = f0[c];
    }
    if (MODE == 1) {
        // Hairer-Wanner initial step as OrdinaryDiffEq codes it (ode_determine_initdt: exponent 1 / alg_order = 1/5 for Tsit5, not 1/(order+1))
        float dt0[CPW], d1[CPW], y1[CPW], f1[CPW];
#pragma unroll
        for (int c = 0; c < CPW; ++c) {
            const float sk = abstol + fabsf(y[c]) * reltol;
            const float a = y[c] / sk, b = f0[c] / sk;
            const float d0 = sqrtf(warp_sum(a * a) * inz);
            d1[c] = sqrtf(warp_sum(b * b) * inz);
            float t = (d0 < 1e-5f || d1[c] < 1e-5f) ? 1e-6f : 0.01f * d0 / d1[c];
            dt0[c] = fminf(t, tf);
            y1[c] = fmaf(dt0[c], f0[c], y[c]);
        }
        rhs_tile<N>(sW, ws, y1, f1, bot, top, kca, pref, lane, nz);
#pragma unroll
        for (int c = 0; c < CPW; ++c) {
            const float sk = abstol + fabsf(y[c]) * reltol;
            const float a = (f1[c] - f0[c]) / sk;
            const float d2 = sqrtf(warp_sum(a * a) * inz) / dt0[c];
            const float m = fmaxf(d1[c], d2);
            const float dt1 = (m <= 1e-15f) ? fmaxf(1e-6f, dt0[c] * 1e-3f) : powf(10.0f, -(2.0f + log10f(m)) / 5.0f);
            float dt = fminf(fminf(100.0f * dt0[c], dt1), tf);
            if (fixed_dt > 0.f) dt = fixed_dt;
            const long long col = col0 + c;
            if (lane == 0 && col < colEnd) dt_out[col] = dt;
        }
    }
}

template <class N, bool WS, int MODE>
__global__ void __launch_bounds__(256, 1)
rhs_init_kernel(const float* __restrict__ wpack, const float* __restrict__ T, const float* __restrict__ colp, float* __restrict__ out,
                float* __restrict__ dt_out, long long B, float pref, int use_kca, float reltol, float abstol, float tf, float fixed_dt, int nz) {
    extern __shared__ __align__(16) float smem[];
    __shared__ uint64_t bar;
    const float* sW = WS ? smem : wpack;
    if (WS) stage_weights(smem, wpack, N::P_TOTAL, &bar);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    float* ws = smem + (WS ? N::P_TOTAL : 0) + warp * N::S_TOTAL;
    const long long ngroups = (B + CPW - 1) / CPW;
    for (long long g = (long long)blockIdx.x * nw + warp; g < ngroups; g += (long long)gridDim.x * nw)
        rhs_init_group<N, MODE>(sW, ws, g * CPW, B, T, colp, out, dt_out, pref, use_kca, reltol, abstol, tf, fixed_dt, lane, nz);
}

// ---------------------------------------------------------------------------------------------------
// Persistent adaptive Tsit5 solve.  grid = #SMs, one CTA per SM, NW warps per CTA, 8 column slots per
// warp.  A slot whose column reaches tf is refilled from a global queue, so step-count divergence
// between columns costs nothing.  No __syncthreads after the weights are staged.
// ---------------------------------------------------------------------------------------------------
struct WarpSlots {
    int col[CPW];
    float t[CPW], dt[CPW], h[CPW], qold[CPW], tnew[CPW], bot[CPW], top[CPW], kca[CPW];
    int si[CPW], nacc[CPW], nrej[CPW], flag[CPW], status[CPW], tlen[CPW];
};
enum : int { FL_ACCEPT = 1, FL_LAST = 2, FL_DONE = 4 };

// adjoint tape entry (one per accepted step): u_n, c1..c4 (Tsit5 dense output in power basis:
// u(t_n + th h) = u_n + h th (c1 + th (c2 + th (c3 + th c4)))), then (t_n, h_n); see nfc_adjoint.cuh
constexpr int TAPE_STRIDE = 5 * NZ + 8;
constexpr int TAPE_T = 5 * NZ, TAPE_H = 5 * NZ + 1;

// c_p = sum_j R[j][p] k_j  (R = OrdinaryDiffEq Tsit5 interpolant coefficients, same numbers as tsit5_btheta)
__device__ __forceinline__ void tsit5_power_basis(const float (&ks)[7][CPW], int c, float& c1, float& c2, float& c3, float& c4) {
    c1 = ks[0][c];
    c2 = -2.763706197274826f * ks[0][c];
    c2 = fmaf(0.13169999999999998f, ks[1][c], c2); c2 = fmaf(3.9302962368947516f, ks[2][c], c2); c2 = fmaf(-12.411077166933676f, ks[3][c], c2);
    c2 = fmaf(37.50931341651104f, ks[4][c], c2); c2 = fmaf(-27.896526289197286f, ks[5][c], c2); c2 = fmaf(1.5f, ks[6][c], c2);
    c3 = 2.9132554618219126f * ks[0][c];
    c3 = fmaf(-0.2234f, ks[1][c], c3); c3 = fmaf(-5.941033872131505f, ks[2][c], c3); c3 = fmaf(30.33818863028232f, ks[3][c], c3);
    c3 = fmaf(-88.1789048947664f, ks[4][c], c3); c3 = fmaf(65.09189467479366f, ks[5][c], c3); c3 = fmaf(-4.0f, ks[6][c], c3);
    c4 = -1.0530884977290216f * ks[0][c];
    c4 = fmaf(0.1017f, ks[1][c], c4); c4 = fmaf(2.490627285651253f, ks[2][c], c4); c4 = fmaf(-16.548102889244902f, ks[3][c], c4);
    c4 = fmaf(47.37952196281928f, ks[4][c], c4); c4 = fmaf(-34.87065786149661f, ks[5][c], c4); c4 = fmaf(2.5f, ks[6][c], c4);
}

// Tsit5 a-matrix rows for stages 2..7 (row 7 == b), zero padded so that one runtime loop serves every stage:
// the Dense-layer code then exists once in the kernel (instruction-cache resident) instead of six times.
static __constant__ float c_tsit5_a[6][6] = {
    {NFC_A21, 0.f, 0.f, 0.f, 0.f, 0.f},
    {NFC_A31, NFC_A32, 0.f, 0.f, 0.f, 0.f},
    {NFC_A41, NFC_A42, NFC_A43, 0.f, 0.f, 0.f},
    {NFC_A51, NFC_A52, NFC_A53, NFC_A54, 0.f, 0.f},
    {NFC_A61, NFC_A62, NFC_A63, NFC_A64, NFC_A65, 0.f},
    {NFC_B1, NFC_B2, NFC_B3, NFC_B4, NFC_B5, NFC_B6}};

// Stages 2..7 of one Tsit5 step for the warp's 8 columns.  On return y == u_new and ks[6] == f(u_new) (FSAL).
// ks[j] of not-yet-computed stages must be finite (they are multiplied by the zero padding).
template <class N>
__device__ __forceinline__ void tsit5_stages(const float* sW, float* ws, const WarpSlots* sl, const float (&u)[CPW], float (&ks)[7][CPW],
                                             float (&y)[CPW], float pref, int lane, int nz = NZ) {
#pragma unroll 1
    for (int s = 0; s < 6; ++s) {
        const float a0 = c_tsit5_a[s][0], a1 = c_tsit5_a[s][1], a2 = c_tsit5_a[s][2], a3 = c_tsit5_a[s][3], a4 = c_tsit5_a[s][4],
                    a5 = c_tsit5_a[s][5];
#pragma unroll
        for (int c = 0; c < CPW; ++c) {
            float acc = a0 * ks[0][c];
            acc = fmaf(a1, ks[1][c], acc); acc = fmaf(a2, ks[2][c], acc); acc = fmaf(a3, ks[3][c], acc);
            acc = fmaf(a4, ks[4][c], acc); acc = fmaf(a5, ks[5][c], acc);
            y[c] = fmaf(sl->h[c], acc, u[c]);
        }
        float f[CPW];
        rhs_tile<N>(sW, ws, y, f, sl->bot, sl->top, sl->kca, pref, lane, nz);
#pragma unroll
        for (int c = 0; c < CPW; ++c) {
            if (s == 0) ks[1][c] = f[c];
            else if (s == 1) ks[2][c] = f[c];
            else if (s == 2) ks[3][c] = f[c];
            else if (s == 3) ks[4][c] = f[c];
            else if (s == 4) ks[5][c] = f[c];
            else ks[6][c] = f[c];
        }
    }
}

// Column queue of one ensemble member (solve_ensemble_kernel): the member's columns [base, base + count) are handed out
// through a shared-memory counter.  The plain solve uses the global queue over all B columns.
struct MemberQueue {
    int* next;        // shared-memory counter
    long long base;
    int count;
};

template <class N, bool RECORD, bool ENS = false>
__device__ __forceinline__ void fill_slot(const SolveParams& P, WarpSlots* sl, int c, float& u, float& k1, double& lsum, int lane,
                                          const MemberQueue& mq = MemberQueue{nullptr, 0, 0}) {
    long long next = 0;
    if (ENS) {
        if (lane == 0) { const int q = atomicAdd(mq.next, 1); next = q < mq.count ? mq.base + q : P.B; }
    } else if (lane == 0) next = (long long)atomicAdd(&P.counters[CTR_QUEUE], 1ULL);
    next = __shfl_sync(FULL, next, 0);
    if (next < P.B) {
        if (!ENS && P.order) next = P.order[next];
        u = P.T0[next * NZ + lane];
        k1 = P.k1_init[next * NZ + lane];
        int si = 0;
        float sse = 0.f;
        while (si < P.n_save && P.saveat[si] <= 0.f) {
            const long long o = (next * P.n_save + si) * NZ + lane;
            if (RECORD) { const float r = u - P.Ttrue[o]; P.resid[o] = r; sse = fmaf(r, r, sse); }
            else if (P.Tout) P.Tout[o] = u;
            ++si;
        }
        if (RECORD && si > 0) lsum += (double)warp_sum(sse);
        if (lane == 0) {
            sl->col[c] = (int)next; sl->t[c] = 0.f; sl->dt[c] = P.dt_init[next]; sl->qold[c] = 1e-4f; sl->si[c] = si;
            sl->nacc[c] = 0; sl->nrej[c] = 0; sl->flag[c] = 0; sl->status[c] = ST_OK; sl->tlen[c] = 0;
            sl->bot[c] = P.colp[next * 3 + 0]; sl->top[c] = P.colp[next * 3 + 1];
            sl->kca[c] = P.use_kca ? P.colp[next * 3 + 2] : 0.f;
        }
    } else {
        u = 0.f; k1 = 0.f;
        if (lane == 0) { sl->col[c] = -1; sl->h[c] = 0.f; sl->dt[c] = 0.f; sl->t[c] = 0.f; sl->bot[c] = 0.f; sl->top[c] = 0.f; sl->kca[c] = 0.f; sl->flag[c] = 0; }
    }
}

// One warp's share of the solve: fill its 8 slots from the queue, step until the queue is empty and the slots have drained.
template <class N, bool RECORD, bool ENS>
__device__ __forceinline__ void solve_warp(const SolveParams& P, const float* sW, float* ws, WarpSlots* sl, int lane, const MemberQueue& mq) {
    float u[CPW], ks[7][CPW], y[CPW];
    double lsum = 0.0;   // this warp's sum of squared residuals (RECORD mode)
    const int nz = P.nz > 0 ? P.nz : NZ;
    const float inz = 1.f / (float)nz;
#pragma unroll
    for (int c = 0; c < CPW; ++c) {
#pragma unroll
        for (int j = 1; j < 7; ++j) ks[j][c] = 0.f;
        fill_slot<N, RECORD, ENS>(P, sl, c, u[c], ks[0][c], lsum, lane, mq);
    }
    __syncwarp();

    while (true) {
        bool any = false;
#pragma unroll
        for (int c = 0; c < CPW; ++c) any |= sl->col[c] >= 0;
        if (!any) break;
        __syncwarp();   // everyone has read col[] before lane<8 rewrites flags
        if (lane < CPW && sl->col[lane] >= 0) {
            const float t = sl->t[lane];
            float h = fminf(sl->dt[lane], P.tf - t);
            int fl = 0;
            if (t + h >= P.tf * (1.0f - 1e-6f)) { h = P.tf - t; fl = FL_LAST; }
            sl->h[lane] = h; sl->flag[lane] = fl;
        }
        __syncwarp();
        tsit5_stages<N>(sW, ws, sl, u, ks, y, P.pref, lane, nz);   // y == u_new, ks[6] == f(u_new) (FSAL)

        // ---- error estimate (RMS of err/(abstol + max(|u|,|unew|) reltol)), every lane gets all 8 ----
        float ee = 0.f;
#pragma unroll
        for (int c = 0; c < CPW; ++c) {
            float e = NFC_BT1 * ks[0][c];
            e = fmaf(NFC_BT2, ks[1][c], e); e = fmaf(NFC_BT3, ks[2][c], e); e = fmaf(NFC_BT4, ks[3][c], e);
            e = fmaf(NFC_BT5, ks[4][c], e); e = fmaf(NFC_BT6, ks[5][c], e); e = fmaf(NFC_BT7, ks[6][c], e);
            e *= sl->h[c];
            const float sc = P.abstol + fmaxf(fabsf(u[c]), fabsf(y[c])) * P.reltol;
            const float r = e / sc;
            const float ss = warp_sum(r * r);
            if (lane == c) ee = sqrtf(ss * inz);
        }
        // ---- PI controller (OrdinaryDiffEq defaults for Tsit5), lane c handles column c ----
        if (lane < CPW && sl->col[lane] >= 0) {
            const float h = sl->h[lane];
            int fl = sl->flag[lane];
            int st = ST_OK;
            bool accept;
            float dtnew;
            if (!isfinite(ee)) { st = ST_NAN; accept = false; dtnew = h; }
            else if (P.fixed_dt > 0.f) { accept = true; dtnew = P.fixed_dt; }
            else {
                const float q11 = powf(ee, 0.14f);
                float q = q11 / powf(sl->qold[lane], 0.08f);
                q = fmaxf(0.1f, fminf(5.0f, q / 0.9f));
                accept = ee <= 1.0f;
                dtnew = accept ? h / q : h / fminf(5.0f, q11 / 0.9f);
            }
            if (accept) {
                const float tn = (fl & FL_LAST) ? P.tf : sl->t[lane] + h;
                sl->tnew[lane] = tn;
                if (P.fixed_dt <= 0.f) sl->qold[lane] = fmaxf(ee, 1e-4f);
                sl->nacc[lane] += 1;
                fl |= FL_ACCEPT;
                if (tn >= P.tf) fl |= FL_DONE;
            } else {
                sl->nrej[lane] += 1;
            }
            sl->dt[lane] = dtnew;
            if (st == ST_OK && !(fl & FL_DONE)) {
                if (dtnew < 1e-12f) st = ST_DT_UNDERFLOW;
                else if (sl->nacc[lane] + sl->nrej[lane] >= P.max_iters) st = ST_MAXITERS;
            }
            if (st != ST_OK) { sl->status[lane] = st; fl |= FL_DONE; }
            sl->flag[lane] = fl;
        }
        __syncwarp();
        // ---- accepted columns: dense output at every save point in (t, t+h], then advance ----
        int mysi = 0;
#pragma unroll
        for (int c = 0; c < CPW; ++c) {
            const int fl = sl->flag[c];
            if (fl & FL_ACCEPT) {
                const long long col = sl->col[c];
                const float t0 = sl->t[c], h = sl->h[c], tn = sl->tnew[c];
                int si = sl->si[c];
                float sse = 0.f;
                float pc1 = 0.f, pc2 = 0.f, pc3 = 0.f, pc4 = 0.f;
                if (RECORD) {
                    tsit5_power_basis(ks, c, pc1, pc2, pc3, pc4);
                    const int n = sl->tlen[c];
                    if (n < tape_capacity(P, col)) {
                        float* e = P.tape + (tape_row0(P, col) + n) * TAPE_STRIDE;
                        e[lane] = u[c]; e[NZ + lane] = pc1; e[2 * NZ + lane] = pc2; e[3 * NZ + lane] = pc3; e[4 * NZ + lane] = pc4;
                        if (lane == 0) { e[TAPE_T] = t0; e[TAPE_H] = h; }
                    }
                }
                while (si < P.n_save) {
                    const float ts = P.saveat[si];
                    if (ts > tn) break;
                    const float th = fminf((ts - t0) / h, 1.0f);
                    const long long o = (col * P.n_save + si) * NZ + lane;
                    if (RECORD) {
                        const float v = fmaf(h, th * fmaf(th, fmaf(th, fmaf(th, pc4, pc3), pc2), pc1), u[c]);
                        const float r = v - P.Ttrue[o]; P.resid[o] = r; sse = fmaf(r, r, sse);
                    } else {
                        float b[7];
                        tsit5_btheta(th, b);
                        float inc = b[0] * ks[0][c];
#pragma unroll
                        for (int j = 1; j < 7; ++j) inc = fmaf(b[j], ks[j][c], inc);
                        if (P.Tout) P.Tout[o] = fmaf(h, inc, u[c]);      // null: counting pass of the compact tape (nfc_adjoint.cuh)
                    }
                    ++si;
                }
                if (RECORD && si != sl->si[c]) lsum += (double)warp_sum(sse);
                if (lane == c) mysi = si;
                u[c] = y[c];
                ks[0][c] = ks[6][c];
            }
        }
        __syncwarp();
        if (lane < CPW && (sl->flag[lane] & FL_ACCEPT)) {
            sl->t[lane] = sl->tnew[lane]; sl->si[lane] = mysi;
            if (RECORD) { if (sl->tlen[lane] >= tape_capacity(P, sl->col[lane])) sl->status[lane] = ST_TAPE_OVERFLOW; sl->tlen[lane] += 1; }
        }
        __syncwarp();
        // ---- retire finished columns and refill their slots from the queue ----
#pragma unroll
        for (int c = 0; c < CPW; ++c) {
            if (sl->flag[c] & FL_DONE) {
                if (lane == 0) {
                    const int col = sl->col[c];
                    if (P.naccept) P.naccept[col] = sl->nacc[c];
                    if (P.nreject) P.nreject[col] = sl->nrej[c];
                    if (P.status) P.status[col] = sl->status[c];
                    if (P.work) P.work[col] = sl->nacc[c] + sl->nrej[c];
                    if (RECORD) P.tape_len[col] = min(sl->tlen[c], tape_capacity(P, col));
                    atomicAdd(&P.counters[CTR_ACCEPT], (unsigned long long)sl->nacc[c]);
                    atomicAdd(&P.counters[CTR_REJECT], (unsigned long long)sl->nrej[c]);
                }
                __syncwarp();
                fill_slot<N, RECORD, ENS>(P, sl, c, u[c], ks[0][c], lsum, lane, mq);
#pragma unroll
                for (int j = 1; j < 7; ++j) ks[j][c] = 0.f;   // stale stages meet zero tableau entries: keep them finite
            }
        }
        __syncwarp();
    }
    if (RECORD && lane == 0) atomicAdd(P.loss_sum, lsum);
}

template <class N, bool WS, int NW, bool RECORD>
__global__ void __launch_bounds__(NW * 32, 1) solve_kernel(const SolveParams P) {
    extern __shared__ __align__(16) float smem[];
    __shared__ uint64_t bar;
    __shared__ WarpSlots slots[NW];
    const float* sW = WS ? smem : P.wpack;
    if (WS) stage_weights(smem, P.wpack, N::P_TOTAL, &bar);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float* ws = smem + (WS ? N::P_TOTAL : 0) + warp * N::S_TOTAL;
    solve_warp<N, RECORD, false>(P, sW, ws, &slots[warp], lane, MemberQueue{nullptr, 0, 0});
}

// ---------------------------------------------------------------------------------------------------
// Ensemble of Chains (src/testing.jl:42-79 compute_nde_solution_history: every saved epoch's NN x every
// simulation, serial there).  Member m has its own packed weights at wpack + m*P_TOTAL and owns the
// columns [m*S, (m+1)*S).  A CTA takes members from a global queue: stage that member's weights,
// k1 / initial dt for its columns, then the same slot loop as solve_kernel over the member's columns.
// ---------------------------------------------------------------------------------------------------
template <class N, bool WS, int NW>
__global__ void __launch_bounds__(NW * 32, 1) solve_ensemble_kernel(const SolveParams P, long long E, int S, float* k1_buf, float* dt_buf) {
    extern __shared__ __align__(16) float smem[];
    __shared__ uint64_t bar;
    __shared__ WarpSlots slots[NW];
    __shared__ long long s_member;
    __shared__ int s_next;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float* ws = smem + (WS ? N::P_TOTAL : 0) + warp * N::S_TOTAL;
    if (threadIdx.x == 0) mbar_init(&bar, 1);
    unsigned parity = 0;
    for (;;) {
        __syncthreads();   // every warp is done with the previous member's weights and queue
        if (threadIdx.x == 0) { s_member = (long long)atomicAdd(&P.counters[CTR_QUEUE], 1ULL); s_next = 0; }
        __syncthreads();
        const long long m = s_member;
        if (m >= E) break;
        const float* gW = P.wpack + m * (long long)N::P_TOTAL;
        const float* sW = WS ? smem : gW;
        if (WS) { restage_weights(smem, gW, N::P_TOTAL, &bar, parity); parity ^= 1u; }
        const long long base = m * S;
        for (int g = warp; g * CPW < S; g += NW)
            rhs_init_group<N, 1>(sW, ws, base + g * CPW, base + S, P.T0, P.colp, k1_buf, dt_buf, P.pref, P.use_kca, P.reltol, P.abstol, P.tf,
                                 P.fixed_dt, lane, P.nz > 0 ? P.nz : NZ);     // run-time level count as in rhs_init_kernel / solve_kernel: one arithmetic, bit for bit
        __syncthreads();   // k1 / dt of the member's columns are visible to the whole CTA
        solve_warp<N, false, true>(P, sW, ws, &slots[warp], lane, MemberQueue{&s_next, base, S});
    }
}

// ---- longest-first schedule from the previous call's per-column step counts (counting sort, 1024 buckets, descending) ----
constexpr int SCHED_BUCKETS = 1024;
static __global__ void sched_hist_kernel(const int* __restrict__ work, int* __restrict__ hist, long long B) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < B) atomicAdd(&hist[min(max(work[i], 0), SCHED_BUCKETS - 1)], 1);
}
static __global__ void sched_scan_kernel(int* __restrict__ hist) {   // in place: hist[b] = number of columns in longer buckets (b' > b)
    __shared__ int s[SCHED_BUCKETS];
    const int t = threadIdx.x;
    s[t] = hist[SCHED_BUCKETS - 1 - t];
    __syncthreads();
    for (int off = 1; off < SCHED_BUCKETS; off <<= 1) { const int v = t >= off ? s[t - off] : 0; __syncthreads(); s[t] += v; __syncthreads(); }
    hist[SCHED_BUCKETS - 1 - t] = t ? s[t - 1] : 0;
}
static __global__ void sched_scatter_kernel(const int* __restrict__ work, int* __restrict__ hist, int* __restrict__ order, long long B) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < B) order[atomicAdd(&hist[min(max(work[i], 0), SCHED_BUCKETS - 1)], 1)] = (int)i;
}

// Flux diagnosis epilogue of solve_nde(ds, ...) (src/solve.jl:32-48): wT[n] = [bottom; NN(T_n); top] - min(0, K_CA dT/dz)
template <class N, bool WS>
__global__ void __launch_bounds__(256, 1)
diagnose_flux_kernel(const float* __restrict__ wpack, const float* __restrict__ Tsol, const float* __restrict__ colp, float* __restrict__ wT,
                     long long B, int n_save, int use_kca, int nz) {
    extern __shared__ __align__(16) float smem[];
    __shared__ uint64_t bar;
    const float* sW = WS ? smem : wpack;
    if (WS) stage_weights(smem, wpack, N::P_TOTAL, &bar);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    float* ws = smem + (WS ? N::P_TOTAL : 0) + warp * N::S_TOTAL;
    const long long nprof = B * n_save, ngroups = (nprof + CPW - 1) / CPW;
    for (long long g = (long long)blockIdx.x * nw + warp; g < ngroups; g += (long long)gridDim.x * nw) {
        float y[CPW];
#pragma unroll
        for (int c = 0; c < CPW; ++c) { const long long p = g * CPW + c; y[c] = p < nprof ? Tsol[p * NZ + lane] : 0.f; }
        nn_forward_tile<N>(sW, ws, y, lane);
        const float* sY = ws + N::S_Y;
#pragma unroll
        for (int c = 0; c < CPW; ++c) {
            const long long p = g * CPW + c;
            if (p >= nprof) continue;
            const long long col = p / n_save;
            const float kca = use_kca ? colp[col * 3 + 2] : 0.f;
            const float ylo = __shfl_up_sync(FULL, y[c], 1);
            float F = colp[col * 3 + 0];
            if (lane > 0) F = sY[c * YS + lane - 1] - fminf(kca * ((y[c] - ylo) * (float)nz), 0.f);
            if (lane < nz) wT[p * (NZ + 1) + lane] = F;
            else if (lane > nz) wT[p * (NZ + 1) + lane] = 0.f;
            if (lane == 0) wT[p * (NZ + 1) + nz] = colp[col * 3 + 1];        // top flux on face nz
        }
        __syncwarp();
    }
}

}  // namespace nfc
