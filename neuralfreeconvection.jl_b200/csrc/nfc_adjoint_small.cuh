// nfc_adjoint_small.cuh -- the backward pass of nfc_loss_grad for SMALL batches: one ocean column per CTA at a time.
//
// The reference trains on 9 simulations (train_free_convection_nde.jl:241, BASELINE config 3).  adjoint_kernel (nfc_adjoint.cuh)
// is a throughput kernel: a warp owns an 8-column register tile whatever the number of live columns, and every stage flushes the
// 24 735 weight-gradient accumulators of its CTA with RED.ADD -- a stage costs ~40 us whether the CTA holds 64 columns or one,
// so a 9-column epoch (290 backward steps x 7 stages, strictly sequential per column) takes 70-80 ms.  Here all 256 threads of a
// CTA work on ONE column: the Dense layers and their transposes are 128-wide matrix-vector products spread over the CTA, and the
// weight gradient is accumulated in REGISTERS (24 735 accumulators = 96 per thread, flushed once per CTA at the end: a first
// version kept them in shared memory and spent a third of every stage on the read-modify-write traffic), so a stage is a handful
// of 0.1-0.3 us phases.  Same algorithm as adjoint_kernel / oracle adjoint_continuous: adaptive backward Tsit5 on lambda
// with tstops at the save times, u(t) from the forward tape, quadrature weight h*b_s folded into the back-propagated vector,
// a rejected step undone by a replay with negative weight.  Default Chain only (Dense 32 -> 128 -> 128 -> 31).
//
// Shared-memory layout: one padded copy of the weights serves both directions -- row stride 129 (33 for the output layer) makes
// "thread = output neuron, loop over inputs" (W[k][n], n across threads) and "thread = input row, loop over outputs" (W[k][n],
// k across threads) both bank-conflict free.  The tape interval the backward solve currently sits in is cached in shared memory
// (7 stages of a step fall into one or two intervals).
#pragma once
#include "nfc_adjoint.cuh"

namespace nfc {
namespace adjs {

using NetD = Net<128, 2, 0>;
constexpr int H = 128, RS = H + 1, RS3 = 33, NT = 256;
constexpr int O_W1 = 0, O_B1 = O_W1 + NZ * RS, O_W2 = O_B1 + H, O_B2 = O_W2 + H * RS, O_W3 = O_B2 + H, O_B3 = O_W3 + H * RS3, W_TOTAL = O_B3 + 32;
constexpr int O_A0 = W_TOTAL, O_A1 = O_A0 + NZ, O_A2 = O_A1 + H, O_G3 = O_A2 + H, O_D2 = O_G3 + NZ, O_D1 = O_D2 + H, O_TAPE = O_D1 + H,
              O_PART = O_TAPE + 5 * NZ, S_TOTAL = O_PART + NT;
constexpr size_t SMEM_BYTES = sizeof(float) * S_TOTAL;        // 104 KB

struct Slot {
    int col, si, nacc, nrej, flag, mode, ipos, tlen, status;
    float t, dt, h, qold, tstop, bot, top, kca, dt_retry, tn, hn;
    int rejrow;
    int icached;          // tape row held in shared memory (-1: none)
};

// theta (Flux layout) index <-> padded shared-memory index
__device__ __forceinline__ int theta_to_smem(int p) {
    if (p < NetD::T_B0) { const int k = p / H, n = p % H; return O_W1 + k * RS + n; }
    if (p < NetD::T_HID) return O_B1 + (p - NetD::T_B0);
    if (p < NetD::T_HID + H * H) { const int q = p - NetD::T_HID, k = q / H, n = q % H; return O_W2 + k * RS + n; }
    if (p < NetD::T_WL) return O_B2 + (p - NetD::T_HID - H * H);
    if (p < NetD::T_BL) { const int q = p - NetD::T_WL, k = q / 31, n = q % 31; return O_W3 + k * RS3 + n; }
    return O_B3 + (p - NetD::T_BL);
}

__global__ void __launch_bounds__(NT, 1) adjoint_small_kernel(const float* __restrict__ theta, const AdjParams P) {
    extern __shared__ __align__(16) float sm[];
    __shared__ Slot S;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < W_TOTAL; i += NT) sm[i] = 0.f;
    __syncthreads();
    for (int p = tid; p < NetD::T_TOTAL; p += NT) sm[theta_to_smem(p)] = theta[p];
    if (tid == 0) S.col = -1;
    float* g = P.gacc + (size_t)blockIdx.x * P.n_params;
    const float c0 = P.pref * (float)NZ;
    // weight-gradient accumulators of this thread:  gw2: layer 2, output n = tid & 127, inputs k = 64 (tid >> 7) + i;  gw1: layer 1, output
    // n = tid & 127, inputs k = 16 (tid >> 7) + i;  gw3: output layer, output n = tid & 31 (< 31), hidden units k = 16 (tid >> 5) + i;
    // biases: gb (threads < 128: b1 and b2 of neuron tid; threads < 31: b3)
    float gw2[H / 2], gw1[NZ / 2], gw3[H / 8], gb1 = 0.f, gb2 = 0.f, gb3 = 0.f;
#pragma unroll
    for (int i = 0; i < H / 2; ++i) gw2[i] = 0.f;
#pragma unroll
    for (int i = 0; i < NZ / 2; ++i) gw1[i] = 0.f;
#pragma unroll
    for (int i = 0; i < H / 8; ++i) gw3[i] = 0.f;
    float lam = 0.f, ls = 0.f, kap[7];
#pragma unroll
    for (int j = 0; j < 7; ++j) kap[j] = 0.f;
    __syncthreads();

    // side launch of a chunk's longest columns: tell the main stream that this kernel holds its SMs (wait_started_kernel, nfc_api.cu)
    if (P.queue_ctr == CTR_ADJ_QUEUE_HEAVY && blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) atomicExch(&P.counters[CTR_HEAVY_STARTED], 1ULL);
    while (true) {
        // ---------------- next column (warp 0) ----------------
        if (warp == 0 && S.col < 0) {
            while (true) {
                long long next = 0;
                if (lane == 0) next = (long long)atomicAdd(&P.counters[P.queue_ctr], 1ULL);
                next = __shfl_sync(FULL, next, 0);
                if (next >= P.B) break;
                if (P.order) next = P.order[next];
                const int len = P.tape_len[next];
                const int st = P.status ? P.status[next] : 0;
                if (st != ST_OK || len <= 0 || P.n_save < 2) continue;
                lam = P.resid[(next * P.n_save + (P.n_save - 1)) * NZ + lane];
                ls = 0.f;
#pragma unroll
                for (int j = 0; j < 7; ++j) kap[j] = 0.f;
                if (lane == 0) {
                    const float* last = P.tape + (tape_row0(P, next) + (len - 1)) * TAPE_STRIDE;
                    S.t = P.tf; S.tstop = P.saveat[P.n_save - 2]; S.si = P.n_save - 2;
                    S.dt = P.fixed_dt > 0.f ? P.fixed_dt : last[TAPE_H]; S.qold = 1e-4f; S.nacc = 0; S.nrej = 0; S.flag = 0; S.rejrow = 0;
                    S.mode = 0; S.ipos = len - 1; S.tlen = len; S.status = ST_OK; S.dt_retry = 0.f;
                    S.bot = P.colp[next * 3 + 0]; S.top = P.colp[next * 3 + 1]; S.kca = P.use_kca ? P.colp[next * 3 + 2] : 0.f;
                    S.icached = -1;
                    S.col = (int)next;
                }
                break;
            }
        }
        __syncthreads();
        if (S.col < 0) break;
        if (tid == 0) {
            if (S.mode == 0) {
                float h = fminf(S.dt, S.t - S.tstop);
                int fl = 0;
                if (S.t - h <= S.tstop + 1e-6f * P.tf && (S.t - h) - S.tstop <= 0.05f * fabsf(h)) { h = S.t - S.tstop; fl = AFL_LAST; } // snap to the save time only when that stretches the step by < 5 %: a wider window re-stretched a rejected step for ever (livelock at t - tstop = 4e-6, 2 of 10^6 columns)
                S.h = h; S.flag = fl;
            } else {
                S.flag &= AFL_LAST;
            }
        }
        __syncthreads();

        // ------------------------------------------------ 7 stages ------------------------------------------------
#pragma unroll 1
        for (int s = 0; s < 7; ++s) {
            float uca = 0.f, winv = 1.f;
            const bool do_dw = s < 6;
            if (warp == 0) {
                const float hc = S.h;
                float acc = c_adj_a[s][0] * kap[0];
                acc = fmaf(c_adj_a[s][1], kap[1], acc); acc = fmaf(c_adj_a[s][2], kap[2], acc); acc = fmaf(c_adj_a[s][3], kap[3], acc);
                acc = fmaf(c_adj_a[s][4], kap[4], acc); acc = fmaf(c_adj_a[s][5], kap[5], acc);
                ls = fmaf(hc, acc, lam);
                // u(t_s) from the forward tape: the interval's five coefficient rows are cached in shared memory
                const float ts = (s == 6 && (S.flag & AFL_LAST)) ? S.tstop : S.t - c_adj_c[s] * hc;
                const float* base = P.tape + tape_row0(P, S.col) * TAPE_STRIDE;
                int n = S.ipos;
                const int len = S.tlen;
                float tn = S.tn, hn = S.hn;
                if (S.icached != n || (n > 0 && ts < tn) || (n + 1 < len && ts > tn + hn)) {
                    tn = base[(long long)n * TAPE_STRIDE + TAPE_T]; hn = base[(long long)n * TAPE_STRIDE + TAPE_H];
                    while (n > 0 && ts < tn) { --n; tn = base[(long long)n * TAPE_STRIDE + TAPE_T]; hn = base[(long long)n * TAPE_STRIDE + TAPE_H]; }
                    while (n + 1 < len && ts > tn + hn) { ++n; tn = base[(long long)n * TAPE_STRIDE + TAPE_T]; hn = base[(long long)n * TAPE_STRIDE + TAPE_H]; }
                    const float* e = base + (long long)n * TAPE_STRIDE;
#pragma unroll
                    for (int j = 0; j < 5; ++j) sm[O_TAPE + j * NZ + lane] = e[j * NZ + lane];
                    __syncwarp();
                    if (lane == 0) { S.ipos = n; S.icached = n; S.tn = tn; S.hn = hn; }
                    __syncwarp();
                }
                const float th = fminf(fmaxf(fast_div(ts - tn, hn), 0.f), 1.f);
                const float u = fmaf(hn, th * fmaf(th, fmaf(th, fmaf(th, sm[O_TAPE + 4 * NZ + lane], sm[O_TAPE + 3 * NZ + lane]), sm[O_TAPE + 2 * NZ + lane]),
                                                   sm[O_TAPE + NZ + lane]), sm[O_TAPE + lane]);
                const float sgn = S.mode ? -1.f : 1.f;
                const float w = do_dw ? sgn * hc * c_adj_b[s] : 1.f;
                winv = fast_div(1.f, w);
                const float lw = w * ls;
                const float lwlo = __shfl_up_sync(FULL, lw, 1);
                const float ulo = __shfl_up_sync(FULL, u, 1);
                const float Fb = lane > 0 ? c0 * (lw - lwlo) : 0.f;                 // dL/dF_k, faces k = 1..31
                const float kk = S.kca * (float)NZ;
                const float sca = (lane > 0 && kk * (u - ulo) < 0.f) ? kk * Fb : 0.f;
                float shi = __shfl_down_sync(FULL, sca, 1);
                if (lane == NZ - 1) shi = 0.f;
                uca = shi - sca;
                sm[O_A0 + lane] = u;
                sm[O_G3 + ((lane + NZ - 1) & (NZ - 1))] = Fb;                       // output neuron n <-> face n + 1 (row 31 = 0)
            }
            __syncthreads();
            // ---- forward: a1 = relu(W1^T a0 + b1): thread pair (n, half) over the 32 inputs
            {
                const int n = tid & (H - 1), half = tid >> 7;
                float acc = 0.f;
#pragma unroll
                for (int k = 0; k < NZ / 2; ++k) { const int kk2 = half * (NZ / 2) + k; acc = fmaf(sm[O_W1 + kk2 * RS + n], sm[O_A0 + kk2], acc); }
                sm[O_PART + tid] = acc;
            }
            __syncthreads();
            if (tid < H) sm[O_A1 + tid] = fmaxf(sm[O_B1 + tid] + (sm[O_PART + tid] + sm[O_PART + H + tid]), 0.f);
            __syncthreads();
            // ---- forward: a2 = relu(W2^T a1 + b2)
            {
                const int n = tid & (H - 1), half = tid >> 7;
                float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll 4
                for (int k = 0; k < H / 2; k += 4) {
                    const int kb = half * (H / 2) + k;
                    a0 = fmaf(sm[O_W2 + (kb + 0) * RS + n], sm[O_A1 + kb + 0], a0);
                    a1 = fmaf(sm[O_W2 + (kb + 1) * RS + n], sm[O_A1 + kb + 1], a1);
                    a2 = fmaf(sm[O_W2 + (kb + 2) * RS + n], sm[O_A1 + kb + 2], a2);
                    a3 = fmaf(sm[O_W2 + (kb + 3) * RS + n], sm[O_A1 + kb + 3], a3);
                }
                sm[O_PART + tid] = (a0 + a1) + (a2 + a3);
            }
            __syncthreads();
            if (tid < H) sm[O_A2 + tid] = fmaxf(sm[O_B2 + tid] + (sm[O_PART + tid] + sm[O_PART + H + tid]), 0.f);
            __syncthreads();
            // ---- output layer backwards: d2[k] = [a2[k] > 0] sum_n W3[k][n] g3[n];  dW3[k][n] += a2[k] g3[n], db3 += g3
            if (tid < H) {
                const int k = tid;
                float acc = 0.f;
#pragma unroll
                for (int n = 0; n < NZ - 1; ++n) acc = fmaf(sm[O_W3 + k * RS3 + n], sm[O_G3 + n], acc);
                sm[O_D2 + k] = sm[O_A2 + k] > 0.f ? acc : 0.f;
            }
            if (do_dw) {
                const int n = tid & (NZ - 1), part = tid >> 5;
                const float gn = sm[O_G3 + n];                      // row 31 is zero
#pragma unroll
                for (int i = 0; i < H / 8; i += 4) {
                    const float4 a = *reinterpret_cast<const float4*>(&sm[O_A2 + part * (H / 8) + i]);
                    gw3[i] = fmaf(a.x, gn, gw3[i]); gw3[i + 1] = fmaf(a.y, gn, gw3[i + 1]); gw3[i + 2] = fmaf(a.z, gn, gw3[i + 2]); gw3[i + 3] = fmaf(a.w, gn, gw3[i + 3]);
                }
                if (part == 0) gb3 += gn;
            }
            __syncthreads();
            // ---- hidden layer backwards: d1[k] = [a1[k] > 0] sum_n W2[k][n] d2[n] (thread pair (k, half));  dW2[k][n] += a1[k] d2[n], db2 += d2
            {
                const int k = tid & (H - 1), half = tid >> 7;
                float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll 4
                for (int n = 0; n < H / 2; n += 4) {
                    const int nb = half * (H / 2) + n;
                    a0 = fmaf(sm[O_W2 + k * RS + nb + 0], sm[O_D2 + nb + 0], a0);
                    a1 = fmaf(sm[O_W2 + k * RS + nb + 1], sm[O_D2 + nb + 1], a1);
                    a2 = fmaf(sm[O_W2 + k * RS + nb + 2], sm[O_D2 + nb + 2], a2);
                    a3 = fmaf(sm[O_W2 + k * RS + nb + 3], sm[O_D2 + nb + 3], a3);
                }
                sm[O_PART + tid] = (a0 + a1) + (a2 + a3);
                if (do_dw) {
                    const float dn = sm[O_D2 + k];                  // this thread's neuron n = k of layer 2; its 64 input rows are 64 half + i
#pragma unroll
                    for (int i = 0; i < H / 2; i += 4) {
                        const float4 a = *reinterpret_cast<const float4*>(&sm[O_A1 + half * (H / 2) + i]);
                        gw2[i] = fmaf(a.x, dn, gw2[i]); gw2[i + 1] = fmaf(a.y, dn, gw2[i + 1]); gw2[i + 2] = fmaf(a.z, dn, gw2[i + 2]); gw2[i + 3] = fmaf(a.w, dn, gw2[i + 3]);
                    }
                    if (half == 0) gb2 += dn;
                }
            }
            __syncthreads();
            if (tid < H) sm[O_D1 + tid] = sm[O_A1 + tid] > 0.f ? (sm[O_PART + tid] + sm[O_PART + H + tid]) : 0.f;
            __syncthreads();
            // ---- first layer backwards: ub[k] = sum_n W1[k][n] d1[n] (k = 0..31, eight threads per row);  dW1[k][n] += a0[k] d1[n], db1 += d1
            {
                const int k = tid & (NZ - 1), part = tid >> 5;            // warp `part` covers neurons [16 part, 16 part + 16)
                float acc = 0.f;
#pragma unroll
                for (int n = 0; n < H / 8; ++n) acc = fmaf(sm[O_W1 + k * RS + part * (H / 8) + n], sm[O_D1 + part * (H / 8) + n], acc);
                sm[O_PART + tid] = acc;
                if (do_dw) {
                    const int n = tid & (H - 1), half = tid >> 7;
                    const float dn = sm[O_D1 + n];
#pragma unroll
                    for (int i = 0; i < NZ / 2; i += 4) {
                        const float4 a = *reinterpret_cast<const float4*>(&sm[O_A0 + half * (NZ / 2) + i]);
                        gw1[i] = fmaf(a.x, dn, gw1[i]); gw1[i + 1] = fmaf(a.y, dn, gw1[i + 1]); gw1[i + 2] = fmaf(a.z, dn, gw1[i + 2]); gw1[i + 3] = fmaf(a.w, dn, gw1[i + 3]);
                    }
                    if (half == 0) gb1 += dn;
                }
            }
            __syncthreads();
            // ---- kappa_s = J^T lam_s
            if (warp == 0) {
                float ub = 0.f;
#pragma unroll
                for (int p = 0; p < 8; ++p) ub += sm[O_PART + 32 * p + lane];
                const float v = (ub + uca) * winv;
                if (s == 0) kap[0] = v; else if (s == 1) kap[1] = v; else if (s == 2) kap[2] = v; else if (s == 3) kap[3] = v;
                else if (s == 4) kap[4] = v; else if (s == 5) kap[5] = v; else kap[6] = v;
            }
            // the next stage's first writes (a0, g3 by warp 0) are separated from this stage's last reads by the barrier above
        }

        // ---------------- error estimate, controller, jump, retire (warp 0) ----------------
        if (warp == 0) {
            float e = NFC_BT1 * kap[0];
            e = fmaf(NFC_BT2, kap[1], e); e = fmaf(NFC_BT3, kap[2], e); e = fmaf(NFC_BT4, kap[3], e);
            e = fmaf(NFC_BT5, kap[4], e); e = fmaf(NFC_BT6, kap[5], e); e = fmaf(NFC_BT7, kap[6], e);
            e *= S.h;
            const float r = fast_div(e, fmaf(fmaxf(fabsf(lam), fabsf(ls)), P.reltol, P.abstol));
            const float ee = fast_sqrt(warp_sum(r * r) * (1.f / NZ));
            __syncwarp();
            if (lane == 0) {
                int fl = S.flag;
                if (S.mode) {                                   // cancel pass done: retry with the reduced step
                    S.mode = 0;
                    S.dt = S.dt_retry;
                } else {
                    const float h = S.h;
                    bool accept;
                    float dtnew;
                    if (P.fixed_dt > 0.f) { accept = true; dtnew = P.fixed_dt; }
                    else if (!isfinite(ee)) { accept = false; dtnew = 0.2f * h; }
                    else {
                        pi_controller(ee, S.qold, h, accept, dtnew);
                    }
                    if (accept) {
                        fl |= AFL_ACCEPT;
                        S.t = (fl & AFL_LAST) ? S.tstop : S.t - h;
                        adj_accept_update(ee, h, S.dt, dtnew, P.fixed_dt <= 0.f, S.qold, S.dt);
                        S.nacc += 1;
                        S.rejrow = 0;
                        if (fl & AFL_LAST) {
                            if (S.si >= 1) fl |= AFL_JUMP; else fl |= AFL_DONE;
                        }
                    } else {
#ifdef NFC_DEBUG_STUCK
                        if (S.nrej > 400 && S.nrej < 412) printf("stuck col %d: t %.9g tstop %.9g h %.9g dt %.9g ee %g dtnew %g si %d ipos %d fl %d\n", S.col, S.t, S.tstop, h, S.dt, ee, dtnew, S.si, S.ipos, fl);
#endif
                        S.nrej += 1;
                        S.mode = 1;                             // undo this step's quadrature contribution next iteration
                        S.dt_retry = dtnew = adj_reject_dt(ee, h, dtnew, S.rejrow);
                        S.rejrow += 1;
                        if (dtnew < 1e-12f || S.nacc + S.nrej >= P.max_iters) S.status = dtnew < 1e-12f ? ST_DT_UNDERFLOW : ST_MAXITERS;
                    }
                    if (S.nacc + S.nrej >= P.max_iters && !(fl & AFL_DONE)) S.status = ST_MAXITERS;
                }
                if (S.status != ST_OK && S.mode == 0) fl |= AFL_DONE;      // a failed column still runs its pending cancel pass first
                S.flag = fl;
            }
            __syncwarp();
            const int fl = S.flag;
            if (fl & AFL_ACCEPT) {
                lam = ls;
                if (fl & AFL_JUMP) lam += P.resid[((long long)S.col * P.n_save + S.si) * NZ + lane];
            }
            __syncwarp();
            if (lane == 0) {
                if (fl & AFL_JUMP) { S.si -= 1; S.tstop = P.saveat[S.si]; }
                if (fl & AFL_DONE) {
                    if (S.status != ST_OK && P.status) P.status[S.col] = S.status;
                    if (P.adj_steps) P.adj_steps[S.col] = S.nacc + S.nrej;
                    atomicAdd(&P.counters[CTR_ADJ_ACCEPT], (unsigned long long)S.nacc);
                    atomicAdd(&P.counters[CTR_ADJ_REJECT], (unsigned long long)S.nrej);
                    S.col = -1;
                }
            }
        }
        __syncthreads();
    }
    // ---------------- flush this thread's gradient accumulators (theta layout; the slice is private to this CTA) ----------------
    {
        const int n = tid & (H - 1), half = tid >> 7;
#pragma unroll
        for (int i = 0; i < H / 2; ++i) g[NetD::T_HID + (half * (H / 2) + i) * H + n] += gw2[i];
#pragma unroll
        for (int i = 0; i < NZ / 2; ++i) g[NetD::T_W0 + (half * (NZ / 2) + i) * H + n] += gw1[i];
        if (half == 0) { g[NetD::T_B0 + n] += gb1; g[NetD::T_HID + H * H + n] += gb2; }
        const int n3 = tid & (NZ - 1), part = tid >> 5;
        if (n3 < NZ - 1) {
#pragma unroll
            for (int i = 0; i < H / 8; ++i) g[NetD::T_WL + (part * (H / 8) + i) * (NZ - 1) + n3] += gw3[i];
            if (part == 0) g[NetD::T_BL + n3] += gb3;
        }
    }
}

}  // namespace adjs
}  // namespace nfc
