// nfc_adjoint.cuh -- loss + gradient (src/training.jl:45-48,70).
//
// The reference differentiates solve(...) with DiffEqSensitivity's continuous adjoint (InterpolatingAdjoint +
// ZygoteVJP, src/solve.jl:5): a second ODE, d(lam)/dt = -J_f(u(t))^T lam, is integrated backwards with the same
// adaptive algorithm, u(t) comes from the forward solution's dense output, lam jumps by dL/du(t_i) at the save
// times and dL/dtheta is the time integral of (df/dtheta)^T lam.  The same algorithm runs here, batched:
//
//   forward  solve_kernel<RECORD>: residuals at the save points + a dense tape (u_n and the Tsit5 interpolant in
//            power basis per accepted step);
//   backward adjoint_kernel: one CTA per SM, 8 warps x 8 column slots, per-slot adaptive Tsit5 on lam with tstops at
//            the save times, slots refilled from a queue.  Per stage every warp recomputes the hidden activations
//            at u(t_s) and back-propagates lam_s through the Chain (warp-private, transposed products read the SAME
//            padded weight copy as the forward pass); the weight-gradient contraction  dW_l += delta_l a_{l-1}^T
//            runs CTA-wide over all 64 columns (quadrature weight h*b_s folded into delta), so every accumulator
//            is touched once per 64 outer products.  Accumulators are per-CTA (no contention), updated with
//            fire-and-forget RED.ADD.F32 in L2; a last kernel reduces them over CTAs in double.
//            A rejected backward step is undone by replaying it with weight -h*b_s ("cancel" pass), so the
//            quadrature never needs per-column gradient storage.
//
// Deliberate differences to DiffEqSensitivity (DESIGN.md "Adjoint"): error control acts on lam only; lam is carried
// un-normalised (jumps = residuals, 2/N applied at the end) so tolerances do not depend on the batch size; no FSAL
// in the backward pass.  oracle/nde_oracle.c:adjoint_continuous is the CPU twin of exactly this algorithm.
#pragma once
#include "nfc_forward.cuh"
#include "nfc_host.h"
#include "nfc_launch.h"

template <class N>
int nfc_launch_record(nfc_handle* h, nfc::SolveParams& P, cudaStream_t st, bool first_chunk);
// counting pass of the compact tape: the forward solve on the kernels (and so the arithmetic) of the RECORD pass, no outputs but P.naccept
template <class N>
int nfc_launch_count(nfc_handle* h, nfc::SolveParams& P, cudaStream_t st);
namespace nfc { struct AdjParams; }
// small batches of the default Chain: one column per CTA (nfc_adjoint_small.cuh); returns false when the batch kernel should run
template <class N>
bool adjoint_small_launch(nfc_handle* h, const nfc::AdjParams& A, int64_t nb, cudaStream_t st);
// the n longest columns of a chunk on the one-column-per-CTA kernel, on the handle's high-priority side stream (default Chain)
int adjoint_heavy_launch(nfc_handle* h, const nfc::AdjParams& A, int n, cudaStream_t st);

namespace nfc {

constexpr int ADJ_NW = 8;   // the CTA-wide dW contraction maps 256 threads onto a 16 x 16 grid of register tiles

struct AdjParams {
    const float* wpack;
    const float* colp;          // [B][3]
    const float* saveat;        // [n_save]
    const float* resid;         // [B][n_save][32]
    const float* tape;          // [B][tape_cap][TAPE_STRIDE], or compact (tape_off, see SolveParams)
    const int* tape_len;        // [B]
    const long long* tape_off;  // [B + 1] or null
    const int* order;           // [B] queue position -> column, longest forward solve first (or null)
    int* status;                // [B] forward status in, adjoint failures merged in
    int* adj_steps;             // [B] attempted backward steps per column (diagnostics) or null
    float* gacc;                // [gridDim.x][n_params]  per-CTA accumulators, Flux theta layout
    unsigned long long* counters;
    long long B;
    int n_save, tape_cap, max_iters, use_kca, n_params;
    int nz;                     // vertical levels in use (<= 32; 0 means 32): adjoint_kernel only
    int queue_ctr;              // index of this launch's queue head in counters[] (CTR_ADJ_QUEUE; the side launch of the longest columns has its own)
    float pref, tf, reltol, abstol, fixed_dt;
};

template <class N>
struct AdjScratch {     // per-warp shared-memory scratch of the backward pass, in floats; every buffer is [rows][8 columns]
    static constexpr int ACT0 = 0;                                  // a_0 = u(t_s)                 [32][8]
    static constexpr int ACT1 = NZ * CPW;                           // a_1 .. a_NHID                [H][8] each
    static constexpr int DL = ACT1 + N::NHID * N::H * CPW;          // delta of the last layer / u-bar staging [32][8]
    static constexpr int TMP = DL + NZ * CPW;                       // delta ping-pong buffer       [H][8]
    static constexpr int TOTAL = TMP + N::H * CPW;
    static __device__ __forceinline__ int act(int l) { return l == 0 ? ACT0 : ACT1 + (l - 1) * N::H * CPW; }
};

struct AdjSlots {
    int col[CPW];
    float t[CPW], dt[CPW], h[CPW], qold[CPW], tstop[CPW], bot[CPW], top[CPW], kca[CPW], dt_retry[CPW];
    int rejrow[CPW];             // rejections in a row (adj_reject_dt)
    float th[CPW], hn[CPW];      // this stage's position inside tape interval ipos: theta in [0, 1] and the interval length
    int si[CPW], nacc[CPW], nrej[CPW], flag[CPW], mode[CPW], ipos[CPW], tlen[CPW], status[CPW];
};
enum : int { AFL_ACCEPT = 1, AFL_LAST = 2, AFL_DONE = 4, AFL_JUMP = 8 };

static __constant__ float c_adj_a[7][6] = {
    {0.f, 0.f, 0.f, 0.f, 0.f, 0.f},
    {NFC_A21, 0.f, 0.f, 0.f, 0.f, 0.f},
    {NFC_A31, NFC_A32, 0.f, 0.f, 0.f, 0.f},
    {NFC_A41, NFC_A42, NFC_A43, 0.f, 0.f, 0.f},
    {NFC_A51, NFC_A52, NFC_A53, NFC_A54, 0.f, 0.f},
    {NFC_A61, NFC_A62, NFC_A63, NFC_A64, NFC_A65, 0.f},
    {NFC_B1, NFC_B2, NFC_B3, NFC_B4, NFC_B5, NFC_B6}};
static __constant__ float c_adj_c[7] = {0.f, 0.161f, 0.327f, 0.9f, 0.9800255409045097f, 1.f, 1.f};
static __constant__ float c_adj_b[7] = {NFC_B1, NFC_B2, NFC_B3, NFC_B4, NFC_B5, NFC_B6, 0.f};

// ---- transposed products on a warp tile: out[k][c] = mask(k,c) * sum_n W[k][pos(n)] din[n][c], lane owns rows k = lane + 32 i ----
// hidden layer: W is [K][RS] with columns in packed order (position NPL*fl + j  <->  neuron fl + 32 j)
template <int KR, int NPL, int RS, bool MASK>
__device__ __forceinline__ void dense_T_hidden(const float* __restrict__ W, const float* __restrict__ din, const float* __restrict__ act,
                                               float* __restrict__ out, int lane, int krows) {
    float acc[KR][CPW];
#pragma unroll
    for (int i = 0; i < KR; ++i)
#pragma unroll
        for (int c = 0; c < CPW; ++c) acc[i][c] = 0.f;
    const bool rowok = lane < krows;           // only relevant for KR == 1 (first layer: krows = IN0 <= 32)
#pragma unroll 2
    for (int fl = 0; fl < 32; ++fl) {
#pragma unroll
        for (int j4 = 0; j4 < NPL / 4; ++j4) {
            float dv[4][CPW];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float* r = din + (fl + 32 * (4 * j4 + j)) * CPW;
                const float4 d0 = *reinterpret_cast<const float4*>(r), d1 = *reinterpret_cast<const float4*>(r + 4);
                dv[j][0] = d0.x; dv[j][1] = d0.y; dv[j][2] = d0.z; dv[j][3] = d0.w; dv[j][4] = d1.x; dv[j][5] = d1.y; dv[j][6] = d1.z; dv[j][7] = d1.w;
            }
#pragma unroll
            for (int i = 0; i < KR; ++i) {
                float4 w = make_float4(0.f, 0.f, 0.f, 0.f);
                if (KR > 1 || rowok) w = *reinterpret_cast<const float4*>(W + (lane + 32 * i) * RS + NPL * fl + 4 * j4);
#pragma unroll
                for (int c = 0; c < CPW; ++c) {
                    float a = acc[i][c];
                    a = fmaf(w.x, dv[0][c], a); a = fmaf(w.y, dv[1][c], a); a = fmaf(w.z, dv[2][c], a); a = fmaf(w.w, dv[3][c], a);
                    acc[i][c] = a;
                }
            }
        }
    }
#pragma unroll
    for (int i = 0; i < KR; ++i) {
        const int k = lane + 32 * i;
        float o[CPW];
        if (MASK) {
            const float4 a0 = *reinterpret_cast<const float4*>(act + k * CPW), a1 = *reinterpret_cast<const float4*>(act + k * CPW + 4);
            o[0] = a0.x > 0.f ? acc[i][0] : 0.f; o[1] = a0.y > 0.f ? acc[i][1] : 0.f; o[2] = a0.z > 0.f ? acc[i][2] : 0.f; o[3] = a0.w > 0.f ? acc[i][3] : 0.f;
            o[4] = a1.x > 0.f ? acc[i][4] : 0.f; o[5] = a1.y > 0.f ? acc[i][5] : 0.f; o[6] = a1.z > 0.f ? acc[i][6] : 0.f; o[7] = a1.w > 0.f ? acc[i][7] : 0.f;
        } else {
#pragma unroll
            for (int c = 0; c < CPW; ++c) o[c] = acc[i][c];
        }
        *reinterpret_cast<float4*>(out + k * CPW) = make_float4(o[0], o[1], o[2], o[3]);
        *reinterpret_cast<float4*>(out + k * CPW + 4) = make_float4(o[4], o[5], o[6], o[7]);
    }
}

// last layer: W is [H][RSL], column n == neuron n (n < 31; columns 31..35 are zero)
template <int KR, int RS>
__device__ __forceinline__ void dense_T_last(const float* __restrict__ W, const float* __restrict__ din, const float* __restrict__ act,
                                             float* __restrict__ out, int lane) {
    float acc[KR][CPW];
#pragma unroll
    for (int i = 0; i < KR; ++i)
#pragma unroll
        for (int c = 0; c < CPW; ++c) acc[i][c] = 0.f;
#pragma unroll 2
    for (int n4 = 0; n4 < 8; ++n4) {
        float dv[4][CPW];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float* r = din + (4 * n4 + j) * CPW;
            const float4 d0 = *reinterpret_cast<const float4*>(r), d1 = *reinterpret_cast<const float4*>(r + 4);
            dv[j][0] = d0.x; dv[j][1] = d0.y; dv[j][2] = d0.z; dv[j][3] = d0.w; dv[j][4] = d1.x; dv[j][5] = d1.y; dv[j][6] = d1.z; dv[j][7] = d1.w;
        }
#pragma unroll
        for (int i = 0; i < KR; ++i) {
            const float4 w = *reinterpret_cast<const float4*>(W + (lane + 32 * i) * RS + 4 * n4);
#pragma unroll
            for (int c = 0; c < CPW; ++c) {
                float a = acc[i][c];
                a = fmaf(w.x, dv[0][c], a); a = fmaf(w.y, dv[1][c], a); a = fmaf(w.z, dv[2][c], a); a = fmaf(w.w, dv[3][c], a);
                acc[i][c] = a;
            }
        }
    }
#pragma unroll
    for (int i = 0; i < KR; ++i) {
        const int k = lane + 32 * i;
        const float4 a0 = *reinterpret_cast<const float4*>(act + k * CPW), a1 = *reinterpret_cast<const float4*>(act + k * CPW + 4);
        *reinterpret_cast<float4*>(out + k * CPW) =
            make_float4(a0.x > 0.f ? acc[i][0] : 0.f, a0.y > 0.f ? acc[i][1] : 0.f, a0.z > 0.f ? acc[i][2] : 0.f, a0.w > 0.f ? acc[i][3] : 0.f);
        *reinterpret_cast<float4*>(out + k * CPW + 4) =
            make_float4(a1.x > 0.f ? acc[i][4] : 0.f, a1.y > 0.f ? acc[i][5] : 0.f, a1.z > 0.f ? acc[i][6] : 0.f, a1.w > 0.f ? acc[i][7] : 0.f);
    }
}

// ---- CTA-wide weight-gradient contraction: gW[k][n] += sum over the CTA's 64 columns of A[k][c] * D[n][c] -----------------
// A / D are the per-warp [rows][8] buffers (warp w's copy at +w*WSTRIDE).  256 threads = 16 x 16 register tiles, rows interleaved
// (k = TK + 16 i, n = TN + 16 j) so that the LDS.128 of a warp hit distinct banks.  gW is this CTA's private accumulator
// (row-major [KROWS][NCOLS] == Flux's column-major out x in), updated with no-return reductions (RED.ADD.F32): a plain
// load-add-store stalled every thread on L2 load latency (30 % of all stall samples in ncu); the addresses are CTA-private.
template <int KROWS, int NCOLS, int WSTRIDE>
__device__ __forceinline__ void dw_gemm(const float* __restrict__ A0, const float* __restrict__ D0, float* __restrict__ gW, float* __restrict__ gb,
                                        const int* __restrict__ warp_active) {
    constexpr int KI = (KROWS + 15) / 16, NJ = (NCOLS + 15) / 16;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int TK = (wid >> 1) * 4 + (lane >> 3), TN = (wid & 1) * 8 + (lane & 7);
    const int hsel = ((lane & 7) >> 2) * 4;       // lanes with tn >= 4 read the second half-row first (bank spreading)
    float acc[KI][NJ], bacc[NJ];
#pragma unroll
    for (int i = 0; i < KI; ++i)
#pragma unroll
        for (int j = 0; j < NJ; ++j) acc[i][j] = 0.f;
#pragma unroll
    for (int j = 0; j < NJ; ++j) bacc[j] = 0.f;
#pragma unroll 1
    for (int w = 0; w < ADJ_NW; ++w) {
        if (!warp_active[w]) continue;               // warps without a live column contribute nothing (and do not fill their buffers)
        const float* A = A0 + w * WSTRIDE;
        const float* D = D0 + w * WSTRIDE;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            const int hd = (half * 4) ^ hsel;        // which 4 of the 8 columns this pass covers (same for both operands)
            float4 a[KI], d[NJ];
#pragma unroll
            for (int i = 0; i < KI; ++i) a[i] = *reinterpret_cast<const float4*>(A + (TK + 16 * i) * CPW + hd);
#pragma unroll
            for (int j = 0; j < NJ; ++j) d[j] = *reinterpret_cast<const float4*>(D + (TN + 16 * j) * CPW + hd);
#pragma unroll
            for (int i = 0; i < KI; ++i)
#pragma unroll
                for (int j = 0; j < NJ; ++j) {
                    float s = acc[i][j];
                    s = fmaf(a[i].x, d[j].x, s); s = fmaf(a[i].y, d[j].y, s); s = fmaf(a[i].z, d[j].z, s); s = fmaf(a[i].w, d[j].w, s);
                    acc[i][j] = s;
                }
#pragma unroll
            for (int j = 0; j < NJ; ++j) bacc[j] += (d[j].x + d[j].y) + (d[j].z + d[j].w);
        }
    }
#pragma unroll
    for (int i = 0; i < KI; ++i) {
        const int k = TK + 16 * i;
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
            const int n = TN + 16 * j;
            if (k < KROWS && n < NCOLS) atomicAdd(gW + k * NCOLS + n, acc[i][j]);   // RED.ADD.F32: fire-and-forget, address private to this CTA
        }
    }
    if (TK == 0) {
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
            const int n = TN + 16 * j;
            if (n < NCOLS) atomicAdd(gb + n, bacc[j]);
        }
    }
}

template <class N>
__device__ __forceinline__ bool adj_fill_slot(const AdjParams& P, AdjSlots* sl, int c, float& lam, int lane) {
    // returns false when the queue is exhausted (slot left inactive)
    while (true) {
        long long next = 0;
        if (lane == 0) next = (long long)atomicAdd(&P.counters[CTR_ADJ_QUEUE], 1ULL);
        next = __shfl_sync(FULL, next, 0);
        if (next >= P.B) {
            lam = 0.f;
            if (lane == 0) { sl->col[c] = -1; sl->h[c] = 0.f; sl->t[c] = 0.f; sl->tstop[c] = 0.f; sl->kca[c] = 0.f; sl->flag[c] = 0; sl->mode[c] = 0; sl->tlen[c] = 0; sl->ipos[c] = 0; }
            return false;
        }
        if (P.order) next = P.order[next];
        const int len = P.tape_len[next];
        const int st = P.status ? P.status[next] : 0;
        if (st != ST_OK || len <= 0 || P.n_save < 2) continue;      // nothing to differentiate (or forward failed: reported in status[])
        lam = P.resid[(next * P.n_save + (P.n_save - 1)) * NZ + lane];      // jump at tf
        if (lane == 0) {
            const float* last = P.tape + (tape_row0(P, next) + (len - 1)) * TAPE_STRIDE;
            sl->col[c] = (int)next; sl->t[c] = P.tf; sl->tstop[c] = P.saveat[P.n_save - 2]; sl->si[c] = P.n_save - 2;
            sl->dt[c] = P.fixed_dt > 0.f ? P.fixed_dt : last[TAPE_H]; sl->qold[c] = 1e-4f; sl->nacc[c] = 0; sl->nrej[c] = 0; sl->flag[c] = 0; sl->rejrow[c] = 0;
            sl->mode[c] = 0; sl->ipos[c] = len - 1; sl->tlen[c] = len; sl->status[c] = ST_OK; sl->dt_retry[c] = 0.f;
            sl->bot[c] = P.colp[next * 3 + 0]; sl->top[c] = P.colp[next * 3 + 1]; sl->kca[c] = P.use_kca ? P.colp[next * 3 + 2] : 0.f;
        }
        return true;
    }
}

template <class N, bool WS>
__global__ void __launch_bounds__(ADJ_NW * 32, 1) adjoint_kernel(const AdjParams P) {
    using S = AdjScratch<N>;
    extern __shared__ __align__(16) float smem[];
    __shared__ uint64_t bar;
    __shared__ AdjSlots slots[ADJ_NW];
    __shared__ int warp_active[ADJ_NW];
    const float* sW = WS ? smem : P.wpack;
    if (WS) stage_weights(smem, P.wpack, N::P_TOTAL, &bar);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float* scratch0 = smem + (WS ? N::P_TOTAL : 0);
    float* ws = scratch0 + warp * S::TOTAL;
    AdjSlots* sl = &slots[warp];
    float* g = P.gacc + (size_t)blockIdx.x * P.n_params;
    const int nz = P.nz > 0 ? P.nz : NZ;
    const float fnz = (float)nz, inz = 1.f / fnz;
    const float c0 = P.pref * fnz;

    float lam[CPW], kap[7][CPW], ls[CPW];
#pragma unroll
    for (int c = 0; c < CPW; ++c) {
#pragma unroll
        for (int j = 0; j < 7; ++j) kap[j][c] = 0.f;
        ls[c] = 0.f;
        adj_fill_slot<N>(P, sl, c, lam[c], lane);
    }
    __syncwarp();

    while (true) {
        bool any = false;
#pragma unroll
        for (int c = 0; c < CPW; ++c) any |= sl->col[c] >= 0;
        if (lane == 0) warp_active[warp] = any ? 1 : 0;
        if (!__syncthreads_or(any ? 1 : 0)) break;
        const bool wact = any;                   // warp-uniform: this warp has at least one live column in this iteration
        if (lane < CPW && sl->col[lane] >= 0) {
            if (sl->mode[lane] == 0) {
                const float t = sl->t[lane], ts = sl->tstop[lane];
                float h = fminf(sl->dt[lane], t - ts);
                int fl = 0;
                if (t - h <= ts + 1e-6f * P.tf && (t - h) - ts <= 0.05f * fabsf(h)) { h = t - ts; fl = AFL_LAST; } // snap to the save time only when that stretches the step by < 5 %: a wider window re-stretched a rejected step for ever (livelock at t - tstop = 4e-6, 2 of 10^6 columns)
                sl->h[lane] = h; sl->flag[lane] = fl;
            } else {
                sl->flag[lane] &= AFL_LAST;          // cancel pass replays the rejected step (same h, same tstop logic)
            }
        }
        __syncwarp();

        // ------------------------------------------------ 7 stages ------------------------------------------------
#pragma unroll 1
        for (int s = 0; s < 7; ++s) {
            const float a0 = c_adj_a[s][0], a1 = c_adj_a[s][1], a2 = c_adj_a[s][2], a3 = c_adj_a[s][3], a4 = c_adj_a[s][4], a5 = c_adj_a[s][5];
            const float cs = c_adj_c[s], bs = c_adj_b[s];
            float us[CPW], uca[CPW], winv[CPW], fb[CPW];
#pragma unroll
            for (int c = 0; c < CPW; ++c) { us[c] = 0.f; uca[c] = 0.f; winv[c] = 1.f; fb[c] = 0.f; }
            if (wact) {
            // u(t_s) from the forward tape.  Lanes 0..7 locate the tape interval of their own column side by side (the search is a
            // chain of dependent global loads), then all 40 coefficient rows of the warp's 8 columns are fetched in one batch: done
            // column by column the exposed latency was 8 % of all stall samples.
            if (lane < CPW) {
                const int col = sl->col[lane];
                float th = 0.f, hn = 0.f;
                if (col >= 0) {
                    const float tcur = sl->t[lane], hc = sl->h[lane];
                    const float ts = (s == 6 && (sl->flag[lane] & AFL_LAST)) ? sl->tstop[lane] : tcur - cs * hc;
                    const float* base = P.tape + tape_row0(P, col) * TAPE_STRIDE;
                    int n = sl->ipos[lane];
                    const int len = sl->tlen[lane];
                    float tn = base[(long long)n * TAPE_STRIDE + TAPE_T];
                    hn = base[(long long)n * TAPE_STRIDE + TAPE_H];
                    while (n > 0 && ts < tn) { --n; tn = base[(long long)n * TAPE_STRIDE + TAPE_T]; hn = base[(long long)n * TAPE_STRIDE + TAPE_H]; }
                    while (n + 1 < len && ts > tn + hn) { ++n; tn = base[(long long)n * TAPE_STRIDE + TAPE_T]; hn = base[(long long)n * TAPE_STRIDE + TAPE_H]; }
                    sl->ipos[lane] = n;
                    th = fminf(fmaxf(fast_div(ts - tn, hn), 0.f), 1.f);
                }
                sl->th[lane] = th; sl->hn[lane] = hn;
            }
            __syncwarp();
            float e0[CPW], e1[CPW], e2[CPW], e3[CPW], e4[CPW];
#pragma unroll
            for (int c = 0; c < CPW; ++c) {
                const int col = sl->col[c];
                const float* e = P.tape + (col >= 0 ? (tape_row0(P, col) + sl->ipos[c]) * TAPE_STRIDE : 0LL);   // idle slot: any valid row
                e0[c] = e[lane]; e1[c] = e[NZ + lane]; e2[c] = e[2 * NZ + lane]; e3[c] = e[3 * NZ + lane]; e4[c] = e[4 * NZ + lane];
            }
#pragma unroll
            for (int c = 0; c < CPW; ++c) {
                float acc = a0 * kap[0][c];
                acc = fmaf(a1, kap[1][c], acc); acc = fmaf(a2, kap[2][c], acc); acc = fmaf(a3, kap[3][c], acc);
                acc = fmaf(a4, kap[4][c], acc); acc = fmaf(a5, kap[5][c], acc);
                const float hc = sl->h[c];
                ls[c] = fmaf(hc, acc, lam[c]);
                const int col = sl->col[c];
                const float th = sl->th[c];
                const float u = col >= 0 ? fmaf(sl->hn[c], th * fmaf(th, fmaf(th, fmaf(th, e4[c], e3[c]), e2[c]), e1[c]), e0[c]) : 0.f;
                us[c] = u;
                // quadrature weight folded into the back-propagated vector (stage 7 carries weight 0: propagate unscaled, skip dW)
                const float sgn = sl->mode[c] ? -1.f : 1.f;
                const float w = (s < 6 && col >= 0) ? sgn * hc * bs : 1.f;
                winv[c] = fast_div(1.f, w);
                const float lw = w * ls[c];
                const float lwlo = __shfl_up_sync(FULL, lw, 1);
                const float ulo = __shfl_up_sync(FULL, u, 1);
                const float Fb = (lane > 0 && lane < nz) ? c0 * (lw - lwlo) : 0.f;  // dL/dF_k, faces k = 1..nz-1 (31)
                const float kk = sl->kca[c] * fnz;
                const float sca = (lane > 0 && lane < nz && kk * (u - ulo) < 0.f) ? kk * Fb : 0.f;
                float shi = __shfl_down_sync(FULL, sca, 1);
                if (lane >= nz - 1) shi = 0.f;
                uca[c] = shi - sca;
                fb[c] = Fb;          // delta of the output layer: row n = k - 1  (row 31 = 0)
            }
            }   // wact
            // write a_0 and the output-layer delta (rows shifted by one: neuron n <-> face n+1)
            if (wact) {
                float* a0p = ws + S::ACT0 + lane * CPW;
                if (N::CONVW == 0) {
                    *reinterpret_cast<float4*>(a0p) = make_float4(us[0], us[1], us[2], us[3]);
                    *reinterpret_cast<float4*>(a0p + 4) = make_float4(us[4], us[5], us[6], us[7]);
                } else {
                    // a_0 = relu(conv(u)): Flux Conv((k,1),1=>1,relu), flipped kernel (train_free_convection_nde.jl:138-153)
                    float x[CPW];
                    const float cb = sW[N::P_CONV + N::CONVW];
#pragma unroll
                    for (int c = 0; c < CPW; ++c) {
                        float acc = 0.f;
#pragma unroll
                        for (int j = 0; j < N::CONVW; ++j) acc = fmaf(sW[N::P_CONV + j], __shfl_down_sync(FULL, us[c], N::CONVW - 1 - j), acc);
                        x[c] = lane < N::IN0 ? fmaxf(acc + cb, 0.f) : 0.f;
                    }
                    *reinterpret_cast<float4*>(a0p) = make_float4(x[0], x[1], x[2], x[3]);
                    *reinterpret_cast<float4*>(a0p + 4) = make_float4(x[4], x[5], x[6], x[7]);
                }
                float* dlp = ws + S::DL + ((lane + NZ - 1) & (NZ - 1)) * CPW;
                *reinterpret_cast<float4*>(dlp) = make_float4(fb[0], fb[1], fb[2], fb[3]);
                *reinterpret_cast<float4*>(dlp + 4) = make_float4(fb[4], fb[5], fb[6], fb[7]);
            }
            __syncwarp();
            // forward: hidden activations a_1 .. a_NHID (the output layer itself is not needed)
            if (wact) {
                dense_hidden<N::IN0, N::NPL, N::RSW, true>(sW + N::P_W0, sW + N::P_B0, ws + S::ACT0, ws + S::act(1), lane);
                __syncwarp();
#pragma unroll
                for (int l = 1; l < N::NHID; ++l) {
                    dense_hidden<N::H, N::NPL, N::RSW, true>(sW + N::P_HID + (l - 1) * N::HIDP_STRIDE, sW + N::P_HID + (l - 1) * N::HIDP_STRIDE + N::H * N::RSW,
                                                             ws + S::act(l), ws + S::act(l + 1), lane);
                    __syncwarp();
                }
            }
            // backward through the output layer
            float* dcur = ws + S::DL;
            float* dnext = ws + S::TMP;
            int dcur_off = S::DL, dnext_off = S::TMP;
            if (wact) dense_T_last<N::NPL, N::RSL>(sW + N::P_WL, dcur, ws + S::act(N::NHID), dnext, lane);
            const bool do_dw = s < 6;
            __syncthreads();
            if (do_dw) dw_gemm<N::H, N::NOUT, S::TOTAL>(scratch0 + S::act(N::NHID), scratch0 + dcur_off, g + N::T_WL, g + N::T_BL, warp_active);
            __syncthreads();
            dcur_off = dnext_off;
            // hidden layers l = NHID-1 .. 1  (Dense_l : a_l -> a_{l+1})
#pragma unroll
            for (int l = N::NHID - 1; l >= 1; --l) {
                dnext_off = (dcur_off == S::TMP) ? S::act(l + 1) : S::TMP;
                if (wact) dense_T_hidden<N::NPL, N::NPL, N::RSW, true>(sW + N::P_HID + (l - 1) * N::HIDP_STRIDE, ws + dcur_off, ws + S::act(l), ws + dnext_off, lane, N::H);
                __syncthreads();
                if (do_dw) dw_gemm<N::H, N::H, S::TOTAL>(scratch0 + S::act(l), scratch0 + dcur_off, g + N::T_HID + (l - 1) * N::HID_STRIDE,
                                                         g + N::T_HID + (l - 1) * N::HID_STRIDE + N::H * N::H, warp_active);
                __syncthreads();
                dcur_off = dnext_off;
            }
            // first layer: u-bar = W0^T delta_1  -> DL buffer (free since the output-layer dW is done)
            if (wact) dense_T_hidden<1, N::NPL, N::RSW, false>(sW + N::P_W0, ws + dcur_off, nullptr, ws + S::DL, lane, N::IN0);
            __syncthreads();
            if (do_dw) dw_gemm<N::IN0, N::H, S::TOTAL>(scratch0 + S::ACT0, scratch0 + dcur_off, g + N::T_W0, g + N::T_B0, warp_active);
            __syncthreads();
            // kappa_s = J^T lam_s
            if (wact) {
                const float4 u0 = *reinterpret_cast<const float4*>(ws + S::DL + lane * CPW), u1 = *reinterpret_cast<const float4*>(ws + S::DL + lane * CPW + 4);
                float ub[CPW] = {u0.x, u0.y, u0.z, u0.w, u1.x, u1.y, u1.z, u1.w};
                if (N::CONVW != 0) {
                    // back through relu(conv(u)): d_i = [pre_i > 0] dx_i ; u-bar_m = sum_j w_j d_{m-(k-1-j)} ; conv weight / bias gradient
                    const float cb = sW[N::P_CONV + N::CONVW];
                    float gw[N::CONVW > 0 ? N::CONVW : 1], gbc = 0.f;
#pragma unroll
                    for (int j = 0; j < N::CONVW; ++j) gw[j] = 0.f;
#pragma unroll
                    for (int c = 0; c < CPW; ++c) {
                        float pre = cb, sh[N::CONVW > 0 ? N::CONVW : 1];
#pragma unroll
                        for (int j = 0; j < N::CONVW; ++j) { sh[j] = __shfl_down_sync(FULL, us[c], N::CONVW - 1 - j); pre = fmaf(sW[N::P_CONV + j], sh[j], pre); }
                        const float d = (lane < N::IN0 && pre > 0.f) ? ub[c] : 0.f;
                        gbc += d;
                        float ubar = 0.f;
#pragma unroll
                        for (int j = 0; j < N::CONVW; ++j) {
                            gw[j] = fmaf(d, sh[j], gw[j]);
                            const float dd = __shfl_up_sync(FULL, d, N::CONVW - 1 - j);
                            if (lane >= N::CONVW - 1 - j) ubar = fmaf(sW[N::P_CONV + j], dd, ubar);
                        }
                        ub[c] = ubar;
                    }
                    if (do_dw) {
#pragma unroll
                        for (int j = 0; j < N::CONVW; ++j) { const float t = warp_sum(gw[j]); if (lane == 0) atomicAdd(g + N::T_CONV + j, t); }
                        const float t = warp_sum(gbc);
                        if (lane == 0) atomicAdd(g + N::T_CONV + N::CONVW, t);
                    }
                }
#pragma unroll
                for (int c = 0; c < CPW; ++c) {
                    const float v = (ub[c] + uca[c]) * winv[c];
                    if (s == 0) kap[0][c] = v;
                    else if (s == 1) kap[1][c] = v;
                    else if (s == 2) kap[2][c] = v;
                    else if (s == 3) kap[3][c] = v;
                    else if (s == 4) kap[4][c] = v;
                    else if (s == 5) kap[5][c] = v;
                    else kap[6][c] = v;
                }
            }
            __syncwarp();
        }
        // after the loop: ls == lam_new (stage-7 input), kap[6] == J^T lam_new

        // ---- error estimate on lam ----
        float ee = 0.f;
#pragma unroll
        for (int c = 0; c < CPW; ++c) {
            float e = NFC_BT1 * kap[0][c];
            e = fmaf(NFC_BT2, kap[1][c], e); e = fmaf(NFC_BT3, kap[2][c], e); e = fmaf(NFC_BT4, kap[3][c], e);
            e = fmaf(NFC_BT5, kap[4][c], e); e = fmaf(NFC_BT6, kap[5][c], e); e = fmaf(NFC_BT7, kap[6][c], e);
            e *= sl->h[c];
            const float r = fast_div(e, fmaf(fmaxf(fabsf(lam[c]), fabsf(ls[c])), P.reltol, P.abstol));      // fast controller arithmetic: one for the three backward kernels
            const float ss = warp_sum(r * r);
            if (lane == c) ee = fast_sqrt(ss * inz);
        }
        if (lane < CPW && sl->col[lane] >= 0) {
            int fl = sl->flag[lane];
            if (sl->mode[lane]) {                              // cancel pass done: retry with the reduced step
                sl->mode[lane] = 0;
                sl->dt[lane] = sl->dt_retry[lane];
            } else {
                const float h = sl->h[lane];
                bool accept;
                float dtnew;
                if (P.fixed_dt > 0.f) { accept = true; dtnew = P.fixed_dt; }
                else if (!isfinite(ee)) { accept = false; dtnew = 0.2f * h; }
                else {
                    pi_controller(ee, sl->qold[lane], h, accept, dtnew);
                }
                if (accept) {
                    fl |= AFL_ACCEPT;
                    sl->t[lane] = (fl & AFL_LAST) ? sl->tstop[lane] : sl->t[lane] - h;
                    adj_accept_update(ee, h, sl->dt[lane], dtnew, P.fixed_dt <= 0.f, sl->qold[lane], sl->dt[lane]);
                    sl->nacc[lane] += 1;
                    sl->rejrow[lane] = 0;
                    if (fl & AFL_LAST) {
                        if (sl->si[lane] >= 1) { fl |= AFL_JUMP; }
                        else fl |= AFL_DONE;
                    }
                } else {
                    sl->nrej[lane] += 1;
                    sl->mode[lane] = 1;                        // undo this step's quadrature contribution next iteration
                    sl->dt_retry[lane] = dtnew = adj_reject_dt(ee, h, dtnew, sl->rejrow[lane]);
                    sl->rejrow[lane] += 1;
                    if (dtnew < 1e-12f || sl->nacc[lane] + sl->nrej[lane] >= P.max_iters) { sl->status[lane] = dtnew < 1e-12f ? ST_DT_UNDERFLOW : ST_MAXITERS; }
                }
                if (sl->nacc[lane] + sl->nrej[lane] >= P.max_iters && !(fl & AFL_DONE)) sl->status[lane] = ST_MAXITERS;
            }
            // a failed column still runs its pending cancel pass (mode == 1) before it is retired
            if (sl->status[lane] != ST_OK && sl->mode[lane] == 0) fl |= AFL_DONE;
            sl->flag[lane] = fl;
        }
        __syncwarp();
#pragma unroll
        for (int c = 0; c < CPW; ++c) {
            const int fl = sl->flag[c];
            if (fl & AFL_ACCEPT) {
                lam[c] = ls[c];
                if (fl & AFL_JUMP) {
                    const int si = sl->si[c];
                    lam[c] += P.resid[((long long)sl->col[c] * P.n_save + si) * NZ + lane];
                }
            }
        }
        __syncwarp();
        if (lane < CPW && (sl->flag[lane] & AFL_JUMP)) {
            const int si = sl->si[lane] - 1;
            sl->si[lane] = si;
            sl->tstop[lane] = P.saveat[si];
        }
        __syncwarp();
#pragma unroll
        for (int c = 0; c < CPW; ++c) {
            if (sl->flag[c] & AFL_DONE) {
                if (lane == 0) {
                    const int col = sl->col[c];
                    if (sl->status[c] != ST_OK && P.status) P.status[col] = sl->status[c];
                    if (P.adj_steps) P.adj_steps[col] = sl->nacc[c] + sl->nrej[c];
                    atomicAdd(&P.counters[CTR_ADJ_ACCEPT], (unsigned long long)sl->nacc[c]);
                    atomicAdd(&P.counters[CTR_ADJ_REJECT], (unsigned long long)sl->nrej[c]);
                }
                __syncwarp();
                adj_fill_slot<N>(P, sl, c, lam[c], lane);
#pragma unroll
                for (int j = 0; j < 7; ++j) kap[j][c] = 0.f;
                ls[c] = 0.f;
            }
        }
        __syncwarp();
    }
}

static __global__ void grad_reduce_kernel(const float* __restrict__ gacc, int ncta, int n_params, double scale, float* __restrict__ grad) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_params) return;
    double s = 0.0;
    for (int b = 0; b < ncta; ++b) s += (double)gacc[(size_t)b * n_params + p];
    grad[p] = (float)(scale * s);
}

static __global__ void count_status_kernel(const int* __restrict__ status, long long n, int which, int* __restrict__ count) {
    int c = 0;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) c += status[i] == which;
    c = __reduce_add_sync(0xffffffffu, c);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(count, c);
}

static __global__ void set_loss_count_kernel(double* loss_sum, double count) { loss_sum[1] = count; }

// ------------------------------------------------ host side ------------------------------------------------
template <class N> struct AdjLaunch { static constexpr bool WS = sizeof(float) * (N::P_TOTAL + ADJ_NW * AdjScratch<N>::TOTAL) + 4096 <= 232448; };
template <class N> constexpr size_t adj_smem_bytes() { return sizeof(float) * ((AdjLaunch<N>::WS ? N::P_TOTAL : 0) + (size_t)ADJ_NW * AdjScratch<N>::TOTAL); }

template <class N>
int adjoint_setup(nfc_handle* h) {
    NFC_CUDA(h, cudaFuncSetAttribute(adjoint_kernel<N, AdjLaunch<N>::WS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)adj_smem_bytes<N>()));
    return 0;
}

template <class N, bool WS, int NW>
int adjoint_loss_grad(nfc_handle* h, const float* T0, const float* colp, float pref, const float* Ttrue, double* loss_sum_dev, float* grad_dev,
                      int* status, int64_t B, int64_t n_total_columns, cudaStream_t st) {
    if (!loss_sum_dev || !grad_dev) return fail(h, -1, "null output");
    const int n_save = h->cfg.n_save, cap = h->cfg.adjoint_max_steps, grid = h->num_sms;
    NFC_CUDA(h, cudaMemsetAsync(loss_sum_dev, 0, 2 * sizeof(double), st));
    NFC_CUDA(h, h->a_gacc.ensure((size_t)2 * grid * h->n_params));        // rows [grid, 2 grid): the side launch of the longest columns
    NFC_CUDA(h, cudaMemsetAsync(h->a_gacc.p, 0, sizeof(float) * (size_t)2 * grid * h->n_params, st));
    if (B == 0) {          // a rank without columns still takes part in the collective
        NFC_CUDA(h, cudaMemsetAsync(grad_dev, 0, sizeof(float) * h->n_params, st));
        if (h->nccl_comm) return comm_allreduce_grad(h, grad_dev, loss_sum_dev, st);
        return 0;
    }
    // chunk the batch so that tape + residuals stay inside the scratch budget (default 48 GiB of the 180 GB, NFC_TAPE_GIB overrides)
    double budget = h->tape_gib > 0 ? h->tape_gib : 48.0;
    if (const char* e = getenv("NFC_TAPE_GIB")) budget = atof(e) > 0 ? atof(e) : budget;
    const double budget_bytes = budget * 1073741824.0;
    const size_t per_col = sizeof(float) * ((size_t)cap * TAPE_STRIDE + (size_t)n_save * NZ);
    // Tape layout.  Dense: every column owns adjoint_max_steps rows (344 KB at the default 512).  Compact: a column owns the rows it
    // needs -- its accepted steps + 1/8 + 6, ~ 87 KB at the measured 110 steps, no step cap, 4 x more columns per chunk.  The counts come
    // from the previous call on the same batch size (a training loop solves the same columns every epoch; a column that outgrows its
    // share is detected after the RECORD pass and the call restarts with exact counts), else from a counting pass: the forward solve on
    // the RECORD pass's kernels without outputs (~ 8 % of a call).  Option tape_mode: 0 = compact when the dense tape of the whole batch
    // would not fit the budget, 1 = dense, 2 = compact.
    const bool compact = h->tape_mode == 2 || (h->tape_mode == 0 && (double)per_col * (double)B > budget_bytes);
    std::vector<int64_t> chunk_begin;              // first column of every chunk, then B
    NFC_CUDA(h, cudaEventRecord(h->ev[2], st));
    int* st_all = status;
    if (!st_all) { NFC_CUDA(h, h->s_status.ensure((size_t)B)); st_all = h->s_status.p; }
    if (compact) { NFC_CUDA(h, h->a_count.ensure((size_t)B)); NFC_CUDA(h, h->a_flag.ensure(1)); }
    if (h->a_count_B != B) { h->a_count_B = B; h->a_count_valid = false; }
    for (int attempt = 0;; ++attempt) {
        const bool from_history = compact && attempt == 0 && h->a_count_valid;
        bool restart = false;
        int64_t max_cols = 0, max_rows = 0;
        chunk_begin.clear();
        h->a_off_host.clear();
        if (attempt > 0) {
            NFC_CUDA(h, cudaMemsetAsync(loss_sum_dev, 0, 2 * sizeof(double), st));
            NFC_CUDA(h, cudaMemsetAsync(h->a_gacc.p, 0, sizeof(float) * (size_t)2 * grid * h->n_params, st));
        }
        if (compact) {
            if (!from_history) {
                SolveParams C{};
                C.T0 = T0; C.colp = colp; C.pref = pref; C.B = B; C.naccept = h->a_count.p;
                if (int rc = nfc_launch_count<N>(h, C, st)) return rc;
            }
            std::vector<int> cnt((size_t)B);
            NFC_CUDA(h, cudaMemcpyAsync(cnt.data(), h->a_count.p, sizeof(int) * (size_t)B, cudaMemcpyDeviceToHost, st));
            NFC_CUDA(h, cudaStreamSynchronize(st));
            // offsets restart at 0 in every chunk; chunk k's nb + 1 offsets start at a_off[chunk_begin[k] + k]
            std::vector<long long>& offs = h->a_off_host;
            offs.reserve((size_t)B + 16);
            const double resid_bytes = sizeof(float) * (double)n_save * NZ, row_bytes = sizeof(float) * (double)TAPE_STRIDE;
            double used = 0.0;
            long long rows = 0;
            chunk_begin.push_back(0);
            offs.push_back(0);
            for (int64_t b = 0; b < B; ++b) {
                const long long c = std::max(cnt[(size_t)b], 0);
                const long long r = std::min<long long>(c + c / 8 + 6, 32767);     // the backward kernels index a column's rows with 15 bits
                const double need = resid_bytes + row_bytes * (double)r;
                if (used + need > budget_bytes && b - chunk_begin.back() >= 1024) {
                    max_cols = std::max(max_cols, b - chunk_begin.back()); max_rows = std::max<int64_t>(max_rows, rows);
                    chunk_begin.push_back(b); offs.push_back(0);
                    used = 0.0; rows = 0;
                }
                used += need; rows += r;
                offs.push_back(rows);
            }
            max_cols = std::max(max_cols, B - chunk_begin.back()); max_rows = std::max<int64_t>(max_rows, rows);
            chunk_begin.push_back(B);
            NFC_CUDA(h, h->a_off.ensure(offs.size()));
            NFC_CUDA(h, cudaMemcpyAsync(h->a_off.p, offs.data(), sizeof(long long) * offs.size(), cudaMemcpyHostToDevice, st));
            NFC_CUDA(h, cudaStreamSynchronize(st));         // offs may be reallocated by the next call while the copy is in flight
        } else {
            int64_t chunk = (int64_t)(budget_bytes / (double)per_col);
            if (chunk < 1024) chunk = 1024;
            if (chunk > B) chunk = B;
            for (int64_t off = 0; off < B; off += chunk) chunk_begin.push_back(off);
            chunk_begin.push_back(B);
            max_cols = chunk; max_rows = chunk * cap;
        }
        h->a_tape_rows = max_rows;
        NFC_CUDA(h, h->a_tape.ensure((size_t)max_rows * TAPE_STRIDE));
        NFC_CUDA(h, h->a_resid.ensure((size_t)max_cols * n_save * NZ));
        NFC_CUDA(h, h->a_tlen.ensure((size_t)max_cols));
        bool first = true;
        for (size_t k = 0; k + 1 < chunk_begin.size(); ++k) {
            const int64_t off = chunk_begin[k], nb = chunk_begin[k + 1] - off;
            const long long* tape_off = compact ? h->a_off.p + off + (int64_t)k : nullptr;
            if (compact && k + 2 == chunk_begin.size()) {     // nfc_get_tape reads the last chunk
                h->a_off_host.erase(h->a_off_host.begin(), h->a_off_host.begin() + (off + (int64_t)k));
            }
            SolveParams P{};
            P.T0 = T0 + off * NZ; P.colp = colp + off * 3; P.pref = pref; P.B = nb; P.status = st_all + off;
            P.Ttrue = Ttrue + off * (int64_t)n_save * NZ; P.resid = h->a_resid.p; P.tape = h->a_tape.p; P.tape_len = h->a_tlen.p;
            P.loss_sum = loss_sum_dev; P.tape_cap = cap; P.tape_off = tape_off;
            if (compact) P.naccept = h->a_count.p + off;       // the next call's tape shares
            h->record_sched = chunk_begin.size() == 2;          // one chunk: longest-first from the previous call's step counts, like nfc_solve
            const int rc_rec = nfc_launch_record<N>(h, P, st, first);
            h->record_sched = false;
            if (rc_rec) return rc_rec;
            if (from_history) {      // shares predicted from the previous call: did a column outgrow its rows?
                int flag = 0;
                NFC_CUDA(h, cudaMemsetAsync(h->a_flag.p, 0, sizeof(int), st));
                count_status_kernel<<<(int)std::min<int64_t>(1024, (nb + 255) / 256), 256, 0, st>>>(st_all + off, nb, ST_TAPE_OVERFLOW, h->a_flag.p);
                NFC_CUDA(h, cudaMemcpyAsync(&flag, h->a_flag.p, sizeof(int), cudaMemcpyDeviceToHost, st));
                NFC_CUDA(h, cudaStreamSynchronize(st));
                h->last_launches += 1;
                if (flag > 0) { restart = true; break; }
            }
            // backward queue, longest first.  Key: the per-column backward step counts of the previous call on the same batch size (a
            // training loop solves the same columns every epoch), else the forward tape length (a forward-stiff column is usually
            // backward-stiff too, but outliers -- 1100 backward steps against a median of 350 -- are only known from history).
            const int* order = nullptr;
            const int* key = nullptr;
            NFC_CUDA(h, h->a_adjsteps.ensure((size_t)B));
            if (h->a_adjhist_B != B) { h->a_adjhist_B = B; h->a_adjhist_valid = false; }
            if (h->use_sched && nb >= 512) {
                key = h->a_adjhist_valid ? h->a_adjsteps.p + off : h->a_tlen.p;
                NFC_CUDA(h, h->d_order.ensure((size_t)nb));
                NFC_CUDA(h, h->d_hist.ensure(SCHED_BUCKETS));
                NFC_CUDA(h, cudaMemsetAsync(h->d_hist.p, 0, sizeof(int) * SCHED_BUCKETS, st));
                const int nblk = (int)((nb + 255) / 256);
                sched_hist_kernel<<<nblk, 256, 0, st>>>(key, h->d_hist.p, nb);
                sched_scan_kernel<<<1, SCHED_BUCKETS, 0, st>>>(h->d_hist.p);
                sched_scatter_kernel<<<nblk, 256, 0, st>>>(key, h->d_hist.p, h->d_order.p, nb);
                NFC_CUDA(h, cudaGetLastError());
                order = h->d_order.p;
                h->last_launches += 3;
            }
            AdjParams A{};
            A.order = order;
            NFC_CUDA(h, cudaMemsetAsync(h->a_adjsteps.p + off, 0, sizeof(int) * nb, st));     // after the order is built (stream order)
            A.adj_steps = h->a_adjsteps.p + off;
            h->a_adjsteps_n = B;
            A.wpack = h->d_wpack.p; A.colp = colp + off * 3; A.saveat = h->d_saveat.p; A.resid = h->a_resid.p; A.tape = h->a_tape.p;
            A.tape_off = tape_off; A.nz = h->nz; A.queue_ctr = CTR_ADJ_QUEUE;
            A.tape_len = h->a_tlen.p; A.status = st_all + off; A.gacc = h->a_gacc.p; A.counters = h->d_counters.p; A.B = nb; A.n_save = n_save;
            A.tape_cap = cap; A.max_iters = h->cfg.max_iters; A.use_kca = h->cfg.nde_type == NFC_NDE_CONVECTIVE_ADJUSTMENT; A.n_params = (int)h->n_params;
            A.pref = pref; A.tf = h->saveat.back(); A.reltol = h->cfg.adjoint_reltol; A.abstol = h->cfg.adjoint_abstol; A.fixed_dt = h->cfg.fixed_dt;
            if (!adjoint_small_launch<N>(h, A, nb, st)) {
                // default Chain, FP16x2 operand scales available: the tensor-core backward kernel (nfc_adjoint_tc.cuh); else FP32 SIMT
                if (h->arch == ARCH_DEFAULT && h->use_tc && h->tc_fp16 && h->adj_tc && cap <= 32767) {
                    // Stragglers: a column with 4 x the median step count keeps its 128-column CTA alive alone (~ 75 us per pass).  Once the
                    // per-column backward step counts are known from the previous call, the adj_heavy longest columns (the head of the
                    // longest-first order) run on the one-column-per-CTA kernel instead (~ 20 us per pass), on a high-priority side stream next
                    // to the tensor-core kernel, whose CTAs on those SMs start when they are free (persistent CTAs, dynamic queue).
                    int n_heavy = 0;
                    if (order && key != h->a_tlen.p && h->adj_heavy > 0 && nb >= 512) n_heavy = (int)std::min<int64_t>(std::min<int64_t>(h->adj_heavy, nb / 16), h->num_sms);
                    if (n_heavy > 0) {
                        if (int rc = adjoint_heavy_launch(h, A, n_heavy, st)) return rc;
                        A.order = order + n_heavy;
                        A.B = nb - n_heavy;
                    }
                    if (int rc = adjoint_tc_launch(h, A, h->atc_scales, A.B, st)) return rc;
                    if (n_heavy > 0) {
                        NFC_CUDA(h, cudaEventRecord(h->ev_heavy[1], h->heavy_stream));
                        NFC_CUDA(h, cudaStreamWaitEvent(st, h->ev_heavy[1], 0));
                        h->last_launches += 1;
                    }
                } else adjoint_kernel<N, AdjLaunch<N>::WS><<<grid, ADJ_NW * 32, adj_smem_bytes<N>(), st>>>(A);
            }
            NFC_CUDA(h, cudaGetLastError());
            h->last_launches += 1;
            first = false;
        }
        if (!restart) break;
        h->a_count_valid = false;
    }   // attempt
    if (compact) h->a_count_valid = true;
    // with a communicator the sums stay un-normalised until the all-reduce has produced the global residual count
    const double ntot = h->nccl_comm ? 1.0 : (double)n_total_columns * n_save * h->nz;
    grad_reduce_kernel<<<((int)h->n_params + 255) / 256, 256, 0, st>>>(h->a_gacc.p, 2 * grid, (int)h->n_params, 2.0 / ntot, grad_dev);
    set_loss_count_kernel<<<1, 1, 0, st>>>(loss_sum_dev, (double)B * n_save * h->nz);
    NFC_CUDA(h, cudaGetLastError());
    if (h->nccl_comm) { if (int rc = comm_allreduce_grad(h, grad_dev, loss_sum_dev, st)) return rc; }
    NFC_CUDA(h, cudaEventRecord(h->ev[3], st));
    h->ev_adj_valid = true;
    h->a_adjhist_valid = true;
    h->last_launches += 2;
    h->last_columns = B;
    return 0;
}

}  // namespace nfc
