// nfc_forward_small.cuh -- the Tsit5 forward solve (and the RECORD pass of nfc_loss_grad) for SMALL batches: one ocean column per CTA.
//
// The reference solves one simulation at a time and trains on nine (train_free_convection_nde.jl:241; BASELINE configs 1 and 3).  The
// tensor-core kernels are throughput kernels: a stage of a 128-column tile costs ~7 K cycles whether the tile holds 128 columns or one, so
// a 9-column solve (~90 strictly sequential steps per column) took 3 ms.  Here all 256 threads of a CTA work on ONE column, as in
// adjoint_small_kernel (nfc_adjoint_small.cuh), whose padded shared-memory weight layout is reused: the Dense layers are 128-wide
// matrix-vector products spread over the CTA (thread pair (neuron, half of the inputs), partial sums through shared memory), warp 0 keeps
// the Runge-Kutta state with lane == vertical level (flux divergence = two shuffles, error norm = a butterfly), FP32 FMAs throughout.
// Same algorithm, controller, dense output, save points, tape and status handling as solve_kernel / solve_tc_kernel (src/solve.jl:4);
// nfc_solve and the RECORD pass use this ONE kernel for the same batch, so mse(solve) == loss of loss_grad also holds here.
// Default Chain only (Dense 32 -> 128 -> 128 -> 31); columns come from the global queue (P.order: longest first), a CTA takes the next
// one when its column retires.
#pragma once
#include "nfc_adjoint_small.cuh"
#include "nfc_tc.cuh"

namespace nfc {
namespace fws {

using adjs::H;
using adjs::NetD;
using adjs::NT;
using adjs::O_A0;
using adjs::O_A1;
using adjs::O_A2;
using adjs::O_B1;
using adjs::O_B2;
using adjs::O_B3;
using adjs::O_G3;
using adjs::O_PART;
using adjs::O_W1;
using adjs::O_W2;
using adjs::O_W3;
using adjs::RS;
using adjs::RS3;
using adjs::W_TOTAL;
constexpr int S_FLOATS = O_PART + NT;                       // weights, a0, a1, a2, nn (O_G3), partial sums
constexpr size_t SMEM_BYTES = sizeof(float) * S_FLOATS;

template <bool RECORD>
__global__ void __launch_bounds__(NT, 1) solve_small_kernel(const float* __restrict__ theta, const SolveParams P) {
    extern __shared__ __align__(16) float sm[];
    __shared__ int s_col;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < W_TOTAL; i += NT) sm[i] = 0.f;
    __syncthreads();
    for (int p = tid; p < NetD::T_TOTAL; p += NT) sm[adjs::theta_to_smem(p)] = theta[p];
    if (tid == 0) s_col = -1;
    // warp 0: state of the column, lane == level
    float u = 0.f, k[7] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, y = 0.f;
    float t = 0.f, dt = 0.f, qold = 1e-4f, bot = 0.f, top = 0.f, kk = 0.f, hs = 0.f;
    int col = -1, si = 0, nacc = 0, nrej = 0, status = ST_OK;
    bool last = false;
    double lsum = 0.0;
    const float cf = -P.pref * (float)NZ;
    __syncthreads();

    while (true) {
        // ---------------- next column (warp 0) ----------------
        if (warp == 0) {
            if (col < 0) {
                long long next = 0;
                if (lane == 0) next = (long long)atomicAdd(&P.counters[CTR_QUEUE], 1ULL);
                next = __shfl_sync(FULL, next, 0);
                if (next < P.B) {
                    if (P.order) next = P.order[next];
                    col = (int)next;
                    u = P.T0[next * NZ + lane];
                    k[0] = P.k1_init[next * NZ + lane];
                    t = 0.f; dt = P.dt_init[next]; qold = 1e-4f; nacc = 0; nrej = 0; status = ST_OK; si = 0;
                    bot = P.colp[next * 3 + 0]; top = P.colp[next * 3 + 1];
                    kk = P.use_kca ? P.colp[next * 3 + 2] * (float)NZ : 0.f;
                    float sse = 0.f;
                    while (si < P.n_save && P.saveat[si] <= 0.f) {
                        const long long o = (next * P.n_save + si) * NZ + lane;
                        if (RECORD) { const float r = u - P.Ttrue[o]; P.resid[o] = r; sse = fmaf(r, r, sse); }
                        else if (P.Tout) __stcs(P.Tout + o, u);
                        ++si;
                    }
                    if (RECORD) lsum += (double)warp_sum(sse);
                }
            }
            if (lane == 0) s_col = col;
            if (col >= 0) {
                hs = fminf(dt, P.tf - t);
                last = false;
                if (t + hs >= P.tf * (1.0f - 1e-6f)) { hs = P.tf - t; last = true; }
            }
        }
        __syncthreads();
        if (s_col < 0) break;

        // ------------------------------------------------ 6 stages ------------------------------------------------
#pragma unroll 1
        for (int s = 0; s < 6; ++s) {
            if (warp == 0) {
                float acc = c_tsit5_a[s][0] * k[0];
                acc = fmaf(c_tsit5_a[s][1], k[1], acc); acc = fmaf(c_tsit5_a[s][2], k[2], acc); acc = fmaf(c_tsit5_a[s][3], k[3], acc);
                acc = fmaf(c_tsit5_a[s][4], k[4], acc); acc = fmaf(c_tsit5_a[s][5], k[5], acc);
                y = fmaf(hs, acc, u);
                sm[O_A0 + lane] = y;
            }
            __syncthreads();
            // ---- a1 = relu(W1^T y + b1): thread pair (n, half) over the 32 inputs
            {
                const int n = tid & (H - 1), half = tid >> 7;
                float acc = 0.f;
#pragma unroll
                for (int i = 0; i < NZ / 2; ++i) { const int kx = half * (NZ / 2) + i; acc = fmaf(sm[O_W1 + kx * RS + n], sm[O_A0 + kx], acc); }
                sm[O_PART + tid] = acc;
            }
            __syncthreads();
            if (tid < H) sm[O_A1 + tid] = fmaxf(sm[O_B1 + tid] + (sm[O_PART + tid] + sm[O_PART + H + tid]), 0.f);
            __syncthreads();
            // ---- a2 = relu(W2^T a1 + b2)
            {
                const int n = tid & (H - 1), half = tid >> 7;
                float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll 4
                for (int i = 0; i < H / 2; i += 4) {
                    const int kb = half * (H / 2) + i;
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
            // ---- NN = W3^T a2 + b3: output n = tid & 31 (row 31 of W3 is zero padding), hidden units 16 (tid >> 5) .. + 15
            {
                const int n = tid & (NZ - 1), part = tid >> 5;
                float acc = 0.f;
#pragma unroll
                for (int i = 0; i < H / 8; ++i) { const int kx = part * (H / 8) + i; acc = fmaf(sm[O_W3 + kx * RS3 + n], sm[O_A2 + kx], acc); }
                sm[O_PART + tid] = acc;
            }
            __syncthreads();
            // ---- f = -pref d/dz ( [bottom; NN; top] - min(0, K_CA dT/dz) ), lane == level (warp 0)
            if (warp == 0) {
                float nn = sm[O_B3 + lane];
#pragma unroll
                for (int p = 0; p < 8; ++p) nn += sm[O_PART + 32 * p + lane];          // NN output `lane` sits on face lane + 1
                const float ylo = __shfl_up_sync(FULL, y, 1);
                const float nnlo = __shfl_up_sync(FULL, nn, 1);
                const float F = lane == 0 ? bot : nnlo - fminf(kk * (y - ylo), 0.f);    // face k = lane
                float Fhi = __shfl_down_sync(FULL, F, 1);
                if (lane == NZ - 1) Fhi = top;
                const float f = (Fhi - F) * cf;
                if (s == 0) k[1] = f; else if (s == 1) k[2] = f; else if (s == 2) k[3] = f; else if (s == 3) k[4] = f; else if (s == 4) k[5] = f;
                else k[6] = f;
            }
            // the next stage's first write (a0 by warp 0) follows this stage's last reads in program order of warp 0; the partial sums it
            // read were written before the barrier above and are rewritten only after the next barrier
        }

        // ---------------- error estimate, controller, output, retire (warp 0) ----------------
        if (warp == 0 && col >= 0) {
            float e = NFC_BT1 * k[0];
            e = fmaf(NFC_BT2, k[1], e); e = fmaf(NFC_BT3, k[2], e); e = fmaf(NFC_BT4, k[3], e); e = fmaf(NFC_BT5, k[4], e); e = fmaf(NFC_BT6, k[5], e);
            const float part = tc::err_term(0.f, e, k[6], hs, u, y, P.abstol, P.reltol);
            const float ee = tc::fast_sqrt(warp_sum(part) * (1.f / NZ));
            bool accept;
            float dtnew;
            if (!isfinite(ee)) { status = ST_NAN; accept = false; dtnew = hs; }
            else if (P.fixed_dt > 0.f) { accept = true; dtnew = P.fixed_dt; }
            else tc::pi_controller(ee, qold, hs, accept, dtnew);
            bool done = false;
            if (accept) {
                const float tn = last ? P.tf : t + hs;
                // dense output in power basis: u(t + th h) = u + h th (k1 + th (c2 + th (c3 + th c4)))
                float c2 = -2.763706197274826f * k[0], c3 = 2.9132554618219126f * k[0], c4 = -1.0530884977290216f * k[0];
                c2 = fmaf(0.13169999999999998f, k[1], c2); c3 = fmaf(-0.2234f, k[1], c3); c4 = fmaf(0.1017f, k[1], c4);
                c2 = fmaf(3.9302962368947516f, k[2], c2); c3 = fmaf(-5.941033872131505f, k[2], c3); c4 = fmaf(2.490627285651253f, k[2], c4);
                c2 = fmaf(-12.411077166933676f, k[3], c2); c3 = fmaf(30.33818863028232f, k[3], c3); c4 = fmaf(-16.548102889244902f, k[3], c4);
                c2 = fmaf(37.50931341651104f, k[4], c2); c3 = fmaf(-88.1789048947664f, k[4], c3); c4 = fmaf(47.37952196281928f, k[4], c4);
                c2 = fmaf(-27.896526289197286f, k[5], c2); c3 = fmaf(65.09189467479366f, k[5], c3); c4 = fmaf(-34.87065786149661f, k[5], c4);
                c2 = fmaf(1.5f, k[6], c2); c3 = fmaf(-4.0f, k[6], c3); c4 = fmaf(2.5f, k[6], c4);
                if (RECORD) {
                    if (nacc < tape_capacity(P, col)) {
                        float* row = P.tape + (tape_row0(P, col) + nacc) * TAPE_STRIDE;
                        row[lane] = u; row[NZ + lane] = k[0]; row[2 * NZ + lane] = c2; row[3 * NZ + lane] = c3; row[4 * NZ + lane] = c4;
                        if (lane == 0) { row[TAPE_T] = t; row[TAPE_H] = hs; }
                    } else status = ST_TAPE_OVERFLOW;
                }
                float sse = 0.f;
                while (si < P.n_save) {
                    const float ts = P.saveat[si];
                    if (ts > tn) break;
                    const float th = fminf(tc::fast_div(ts - t, hs), 1.0f);
                    const float v = fmaf(hs, th * fmaf(th, fmaf(th, fmaf(th, c4, c3), c2), k[0]), u);
                    const long long o = ((long long)col * P.n_save + si) * NZ + lane;
                    if (RECORD) { const float r = v - P.Ttrue[o]; P.resid[o] = r; sse = fmaf(r, r, sse); }
                    else if (P.Tout) __stcs(P.Tout + o, v);
                    ++si;
                }
                if (RECORD) lsum += (double)warp_sum(sse);
                u = y; k[0] = k[6]; t = tn;
                if (P.fixed_dt <= 0.f) qold = fmaxf(ee, 1e-4f);
                ++nacc;
                if (tn >= P.tf) done = true;
            } else {
                ++nrej;
            }
            dt = dtnew;
            if (status == ST_OK && !done) {
                if (dtnew < 1e-12f) status = ST_DT_UNDERFLOW;
                else if (nacc + nrej >= P.max_iters) status = ST_MAXITERS;
            }
            if (status != ST_OK) done = true;
            if (done) {
                if (lane == 0) {
                    if (P.naccept) P.naccept[col] = nacc;
                    if (P.nreject) P.nreject[col] = nrej;
                    if (P.status) P.status[col] = status;
                    if (P.work) P.work[col] = nacc + nrej;
                    if (RECORD) P.tape_len[col] = min(nacc, tape_capacity(P, col));
                    atomicAdd(&P.counters[CTR_ACCEPT], (unsigned long long)nacc);
                    atomicAdd(&P.counters[CTR_REJECT], (unsigned long long)nrej);
                }
                col = -1;
            }
        }
        __syncthreads();          // warp 0 has read the last partial sums; s_col is rewritten at the top
    }
    if (RECORD && warp == 0 && lane == 0) atomicAdd(P.loss_sum, lsum);
}

}  // namespace fws
}  // namespace nfc
