// nfc_adjoint_tc.cuh -- the backward pass of nfc_loss_grad (src/training.jl:45-48,70) on the tensor cores (tcgen05 / TMEM).
//
// Same algorithm as adjoint_kernel (nfc_adjoint.cuh) and oracle/nde_oracle.c:adjoint_continuous: adaptive backward Tsit5 on lambda with
// tstops at the save times, u(t) from the forward tape, quadrature weight h*b_s folded into the back-propagated vector, a rejected
// step undone by a replay with negative weight.  What changes is where the arithmetic runs.  One CTA owns 128 column slots = the M
// dimension of tcgen05.mma (TMEM lane == ocean column); per Runge-Kutta stage the elected MMA thread issues
//
//   forward recompute   z1 = u W1^T            z2 = a1 W2^T                        TS form, A (activations) in tensor memory,
//   transposed chain    g2 = d3 W3             g1 = d2 W2         ubar = d1 W1     B = the SAME K-major FP16 hi/lo weight image as the
//                                                                                   forward solve; the transposed products read it with
//                                                                                   the MN-major flag (a core matrix of 8 rows x 16 B is
//                                                                                   its own transpose target: only LBO and SBO swap)
//   weight gradient     dW3^T += a2^T d3       dW2 += d2^T [1 a1]  dW1 += d1^T [u 1]   SS form with K = the 128 ocean columns: both
//                                                                                   operands are staged by the epilogue threads as BF16 in
//                                                                                   MN-major layout, the accumulators (224 TMEM columns)
//                                                                                   collect one Runge-Kutta step and are flushed in stage 7; the constant-one rows
//                                                                                   make the bias gradients fall out of the same MMAs.
//
// Arithmetic: the chain (forward and transposed) is FP32-faithful like the forward solve (FP16 hi + lo, three products per k-step,
// exact power-of-two operand scales; the back-propagated vector is normalised per column by a power of two because lambda is carried
// un-normalised and spans many decades).  The weight-gradient contraction is a quadrature SUM over columns x steps x stages that does not
// feed back into the step-size control, so its operands are single BF16 (fp32 range, 8 bits): rounding errors are unbiased and average
// out over the > 10^6 terms of every accumulator (measured against the FP32 kernel in tests/test_gpu_adjoint_tc.py).
//
// Thread roles as in nfc_tc.cuh: warps 0..15 = four threads per column (quarter q owns levels [8q, 8q+8) and hidden units [32q, 32q+32)),
// warp 16 = MMA issuer.  Stage vectors kappa_1..kappa_6: one in tensor memory, one in shared memory, four in registers.
#pragma once
#include "nfc_adjoint.cuh"
#include "nfc_tc.cuh"

namespace nfc {
namespace atc {

using namespace tc;

// tensor-memory columns
constexpr int TA_ACC2 = 0;      // [d2 unit][db2 (16, all equal) | dW2 (128: a1 unit)]
constexpr int TA_ACC1 = 144;    // [d1 unit][dW1 (32: level) | db1 (16)]
constexpr int TA_ACC3 = 192;    // [a2 unit][dW3^T (32: output neuron)]
constexpr int TA_D = 224;       // chain accumulator, 128 columns
constexpr int TA_A = 352;       // chain A operand: hi 64 | lo 64
constexpr int TA_K1 = 480;      // kappa_1, 32 columns
// shared memory (bytes).  A staging "chunk" = 8 units x 128 columns of BF16 = 2048 B in MN-major core matrices: [column][8 units]
constexpr int CH = 2048;
constexpr int SO_U = 100352;                 // u          4 chunks  } B operand of dW1: [u | 1]   (N = 48)
constexpr int SO_ONES = SO_U + 4 * CH;       // ones       2 chunks  } B operand of dW2: [1 | a1]  (N = 144)
constexpr int SO_A1 = SO_ONES + 2 * CH;      // a1        16 chunks  }
constexpr int SO_X = SO_A1 + 16 * CH;        // a2, then d2, then d1 (A operands of the three contractions), 16 chunks
constexpr int SO_D3 = SO_X + 16 * CH;        // d3         4 chunks  (B operand of dW3^T, N = 32)
constexpr int SO_UB = SO_D3 + 4 * CH;        // u(t_s) of the NEXT stage, fp32 [column][32 levels], 16-byte groups XOR-swizzled by the column
constexpr int SO_ST = SO_UB + 16384;
static_assert(IMG16_BYTES <= SO_U, "weight image overlaps the staging buffers");

struct AtcShared {
    uint64_t bar_w, bar_a, bar_a1, bar_d, bar_x;     // bar_a / bar_a1: first / second half of an A operand published (even / odd k-steps)
    uint32_t tmem_base;
    int stop, any;
    float xu[4][2][TILE], xl[4][2][TILE], xm[4][TILE];      // boundary levels of u and w*lambda_s, max |w lambda_s| of every quarter
    float part[4][TILE];
    int bad[TILE];
    // per-column slot state, owned by the quarter-0 thread of the column
    int col[TILE], si[TILE], nacc[TILE], nrej[TILE], mode[TILE], tlen[TILE], status[TILE], flag[TILE], jsi[TILE], ipos[TILE];
    float t[TILE], dt[TILE], h[TILE], qold[TILE], tstop[TILE], dt_retry[TILE], kk[TILE];
    int rejrow[TILE];                                       // rejections in a row (adj_reject_dt)
    short ivp[7][TILE];                                     // tape interval of every stage of the current step
    long long row0[TILE];                                   // first tape row of the slot's column (dense or compact tape)
};
constexpr int ATC_SMEM = SO_ST + (int)sizeof(AtcShared);
static_assert(ATC_SMEM <= 232448, "shared memory budget");
enum : int { AFL_NEWCOL = 16 };

__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | (1ull << 46);
}
// kind::f16 instruction descriptor: fmt 0 = F16, 1 = BF16; C = F32; major bits: 0 = K-major, 1 = MN-major
__host__ __device__ constexpr uint32_t idesc(int fmt, int a_mn, int b_mn, int N) {
    return (1u << 4) | ((uint32_t)fmt << 7) | ((uint32_t)fmt << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
template <bool ACC>
__device__ __forceinline__ void mma_ss2(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t id) {
    if (ACC)
        asm volatile("{\n\t.reg .b64 ad, bd;\n\t.reg .pred p;\n\tsetp.eq.u32 p, 1, 1;\n\tmov.b64 ad, {%1, %2};\n\tmov.b64 bd, {%3, %4};\n\t"
                     "tcgen05.mma.cta_group::1.kind::f16 [%0], ad, bd, %5, p;\n\t}" ::"r"(d_tmem), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(id) : "memory");
    else
        asm volatile("{\n\t.reg .b64 ad, bd;\n\t.reg .pred p;\n\tsetp.ne.u32 p, 1, 1;\n\tmov.b64 ad, {%1, %2};\n\tmov.b64 bd, {%3, %4};\n\t"
                     "tcgen05.mma.cta_group::1.kind::f16 [%0], ad, bd, %5, p;\n\t}" ::"r"(d_tmem), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(id) : "memory");
}
__device__ __forceinline__ void mma_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t id, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem), "l"(a_desc),
                 "l"(b_desc), "r"(id), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// TS-form product over NK k-steps: D (+)= A[cells a + 8 j] x B[desc + kadv j]
template <int NK, bool FIRST>
__device__ __forceinline__ void issue_ts(uint32_t d, uint32_t a, uint32_t blo, uint32_t bhi, uint32_t kadv, uint32_t id) {
#pragma unroll
    for (int j = 0; j < NK; ++j) {
        if (FIRST && j == 0) mma_ts2<false>(d, a + 8 * j, blo + kadv * j, bhi, id);
        else mma_ts2<true>(d, a + 8 * j, blo + kadv * j, bhi, id);
    }
}
// the k-steps of one parity (HALF 0: even = the first halves of the quarters' operand slices, published first)
template <int NK, int HALF, bool FIRST>
__device__ __forceinline__ void issue_ts_half(uint32_t d, uint32_t a, uint32_t blo, uint32_t bhi, uint32_t kadv, uint32_t id) {
#pragma unroll
    for (int j = HALF; j < NK; j += 2) {
        if (FIRST && j == 0) mma_ts2<false>(d, a + 8 * j, blo + kadv * j, bhi, id);
        else mma_ts2<true>(d, a + 8 * j, blo + kadv * j, bhi, id);
    }
}
// SS-form contraction over the 128 columns (8 k-steps of 16 columns = 256 B in both staging buffers).  The accumulators collect ONE
// Runge-Kutta step (first = stage 0 overwrites) and are flushed to the CTA's fp32 gradient row in stage 7: the tensor core adds with
// truncation, and an accumulator that lived for the whole kernel (17 000 additions of increments far below its own magnitude) drifted
// by 1e-3 relative -- growing with the number of steps, which is how it was found (tight tolerances made the gradient WORSE).
__device__ __forceinline__ void issue_dw(uint32_t d, uint32_t a_addr, uint32_t b_addr, uint32_t id, bool first) {
    const uint64_t ad = make_desc(a_addr, 128, CH), bd = make_desc(b_addr, 128, CH);
    const uint32_t alo = (uint32_t)ad, ahi = (uint32_t)(ad >> 32), blo = (uint32_t)bd, bhi = (uint32_t)(bd >> 32);
    mma_ss(d, ad, bd, id, first ? 0u : 1u);
#pragma unroll
    for (int j = 1; j < 8; ++j) mma_ss2<true>(d, alo + 16 * j, ahi, blo + 16 * j, bhi, id);
}

// ---- MMA issuer: per stage five batches, each released by the 512 column threads through bar_a -----------------------------------------
__device__ __forceinline__ void atc_mma_loop(AtcShared* S, uint32_t sbase) {
    int pa = 0;
    constexpr uint32_t ID_F = idesc(0, 0, 0, 128);          // forward layers: K-major weights
    constexpr uint32_t ID_T128 = idesc(0, 0, 1, 128);       // transposed products: the same image, MN-major
    constexpr uint32_t ID_T64 = idesc(0, 0, 1, 64);
    constexpr uint32_t ID_W32 = idesc(1, 1, 1, 32), ID_W144 = idesc(1, 1, 1, 144), ID_W48 = idesc(1, 1, 1, 48);
    while (true) {
#pragma unroll 1
        for (int s = 0; s < 7; ++s) {
            const bool dw = s < 6;
            uint32_t tb, img;
            // (1) z1 = u W1^T
            mbar_wait(&S->bar_a, pa);
            if (s == 0 && *reinterpret_cast<volatile int*>(&S->stop)) return;
            mbar_wait(&S->bar_a1, pa); pa ^= 1;
            fence_after_sync();
            tb = S->tmem_base; img = sbase; asm volatile("" : "+r"(tb), "+r"(img));
            {
                const uint64_t b = make_desc(img + OFF16_W1, 128 * 16, 128);
                const uint32_t lo = (uint32_t)b, hi = (uint32_t)(b >> 32);
                issue_ts<2, true>(tb + TA_D, tb + TA_A + 64, lo, hi, 256, ID_F);                        // a_lo W_hi
                issue_ts<2, false>(tb + TA_D, tb + TA_A, lo + (W1H_BYTES >> 4), hi, 256, ID_F);         // a_hi W_lo
                issue_ts<2, false>(tb + TA_D, tb + TA_A, lo, hi, 256, ID_F);                            // a_hi W_hi
            }
            mma_commit(&S->bar_d);
            // (2) z2 = a1 W2^T: the even k-steps start as soon as the first halves of the operand are published
            mbar_wait(&S->bar_a, pa);
            fence_after_sync();
            tb = S->tmem_base; img = sbase; asm volatile("" : "+r"(tb), "+r"(img));
            {
                const uint64_t b = make_desc(img + OFF16_W2, 128 * 16, 128);
                const uint32_t lo = (uint32_t)b, hi = (uint32_t)(b >> 32);
                issue_ts_half<8, 0, true>(tb + TA_D, tb + TA_A + 64, lo, hi, 256, ID_F);
                issue_ts_half<8, 0, false>(tb + TA_D, tb + TA_A, lo + (W2H_BYTES >> 4), hi, 256, ID_F);
                issue_ts_half<8, 0, false>(tb + TA_D, tb + TA_A, lo, hi, 256, ID_F);
            }
            mbar_wait(&S->bar_a1, pa); pa ^= 1;
            fence_after_sync();
            tb = S->tmem_base; img = sbase; asm volatile("" : "+r"(tb), "+r"(img));
            {
                const uint64_t b = make_desc(img + OFF16_W2, 128 * 16, 128);
                const uint32_t lo = (uint32_t)b, hi = (uint32_t)(b >> 32);
                issue_ts_half<8, 1, false>(tb + TA_D, tb + TA_A + 64, lo, hi, 256, ID_F);
                issue_ts_half<8, 1, false>(tb + TA_D, tb + TA_A, lo + (W2H_BYTES >> 4), hi, 256, ID_F);
                issue_ts_half<8, 1, false>(tb + TA_D, tb + TA_A, lo, hi, 256, ID_F);
            }
            mma_commit(&S->bar_d);
            // (3) g2 = d3 W3 (image: [hi rows 0..31; lo rows 32..63] x K = 128, i.e. transposed: k' 0..31 hi, 32..63 lo), then dW3^T += a2^T d3
            mbar_wait(&S->bar_a, pa);
            mbar_wait(&S->bar_a1, pa); pa ^= 1;
            fence_after_sync();
            tb = S->tmem_base; img = sbase; asm volatile("" : "+r"(tb), "+r"(img));
            {
                const uint64_t b = make_desc(img + OFF16_W3, 128, 64 * 16);
                const uint32_t lo = (uint32_t)b, hi = (uint32_t)(b >> 32);
                issue_ts<2, true>(tb + TA_D, tb + TA_A + 64, lo, hi, 16, ID_T128);                      // d_lo W_hi
                issue_ts<2, false>(tb + TA_D, tb + TA_A, lo + 32, hi, 16, ID_T128);                     // d_hi W_lo (k' 32..63: +512 B)
                issue_ts<2, false>(tb + TA_D, tb + TA_A, lo, hi, 16, ID_T128);                          // d_hi W_hi
            }
            mma_commit(&S->bar_d);
            if (dw) {
                asm volatile("" : "+r"(tb), "+r"(img));
                issue_dw(tb + TA_ACC3, img + SO_X, img + SO_D3, ID_W32, s == 0);
                mma_commit(&S->bar_x);
            }
            // (4) g1 = d2 W2, then dW2 += d2^T [1 a1]
            mbar_wait(&S->bar_a, pa);
            fence_after_sync();
            tb = S->tmem_base; img = sbase; asm volatile("" : "+r"(tb), "+r"(img));
            {
                const uint64_t b = make_desc(img + OFF16_W2, 128, 128 * 16);
                const uint32_t lo = (uint32_t)b, hi = (uint32_t)(b >> 32);
                issue_ts_half<8, 0, true>(tb + TA_D, tb + TA_A + 64, lo, hi, 16, ID_T128);
                issue_ts_half<8, 0, false>(tb + TA_D, tb + TA_A, lo + (W2H_BYTES >> 4), hi, 16, ID_T128);
                issue_ts_half<8, 0, false>(tb + TA_D, tb + TA_A, lo, hi, 16, ID_T128);
            }
            mbar_wait(&S->bar_a1, pa); pa ^= 1;
            fence_after_sync();
            tb = S->tmem_base; img = sbase; asm volatile("" : "+r"(tb), "+r"(img));
            {
                const uint64_t b = make_desc(img + OFF16_W2, 128, 128 * 16);
                const uint32_t lo = (uint32_t)b, hi = (uint32_t)(b >> 32);
                issue_ts_half<8, 1, false>(tb + TA_D, tb + TA_A + 64, lo, hi, 16, ID_T128);
                issue_ts_half<8, 1, false>(tb + TA_D, tb + TA_A, lo + (W2H_BYTES >> 4), hi, 16, ID_T128);
                issue_ts_half<8, 1, false>(tb + TA_D, tb + TA_A, lo, hi, 16, ID_T128);
            }
            mma_commit(&S->bar_d);
            if (dw) {
                asm volatile("" : "+r"(tb), "+r"(img));
                issue_dw(tb + TA_ACC2, img + SO_X, img + SO_ONES, ID_W144, s == 0);
                mma_commit(&S->bar_x);
            }
            // (5) ubar = d1 W1 (hi and lo images are adjacent along the transposed N: one N = 64 operand), then dW1 += d1^T [u 1]
            mbar_wait(&S->bar_a, pa);
            fence_after_sync();
            tb = S->tmem_base; img = sbase; asm volatile("" : "+r"(tb), "+r"(img));
            {
                const uint64_t b = make_desc(img + OFF16_W1, 128, 128 * 16);
                const uint32_t lo = (uint32_t)b, hi = (uint32_t)(b >> 32);
                issue_ts_half<8, 0, true>(tb + TA_D, tb + TA_A + 64, lo, hi, 16, ID_T64);
                issue_ts_half<8, 0, false>(tb + TA_D, tb + TA_A, lo, hi, 16, ID_T64);
            }
            mbar_wait(&S->bar_a1, pa); pa ^= 1;
            fence_after_sync();
            tb = S->tmem_base; img = sbase; asm volatile("" : "+r"(tb), "+r"(img));
            {
                const uint64_t b = make_desc(img + OFF16_W1, 128, 128 * 16);
                const uint32_t lo = (uint32_t)b, hi = (uint32_t)(b >> 32);
                issue_ts_half<8, 1, false>(tb + TA_D, tb + TA_A + 64, lo, hi, 16, ID_T64);
                issue_ts_half<8, 1, false>(tb + TA_D, tb + TA_A, lo, hi, 16, ID_T64);
            }
            mma_commit(&S->bar_d);
            if (dw) {
                asm volatile("" : "+r"(tb), "+r"(img));
                issue_dw(tb + TA_ACC1, img + SO_X, img + SO_U, ID_W48, s == 0);
                mma_commit(&S->bar_x);
            }
        }
    }
}

// (max(lo,0), max(hi,0)) -> packed bf16
__device__ __forceinline__ uint32_t pack_bf16_relu(float lo, float hi) {
    uint32_t r;
    asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}
// packed bf16 pair times a packed power of two (exact): the per-column scale of the staged deltas costs one instruction per pair
__device__ __forceinline__ uint32_t mul_bf16x2(uint32_t a, uint32_t b) {
    uint32_t r;
    asm("mul.rn.bf16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
    return r;
}
// relu mask, one instruction per activation: shift the sign bit of the pre-activation into m (MSB first: after 32 calls element e sits at
// bit 31 - e and m holds "v < 0"; the caller inverts it once).  v == +0 counts as active, harmlessly: its delta meets a zero activation
__device__ __forceinline__ void mask_push(uint32_t& m, float v) { m = __funnelshift_l(__float_as_uint(v), m, 1); }
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}

// sum of v[0..7] over the 32 lanes of the warp; lane l returns the total of element 4 b4 + 2 b3 + b2 (bits of l), replicated over l & 3
__device__ __forceinline__ float warp_sum8(const float (&v)[8], int lane) {
    float w4[4], w2[2];
    const bool u4 = lane & 16, u3 = lane & 8, u2 = lane & 4;
#pragma unroll
    for (int i = 0; i < 4; ++i) { const float snd = u4 ? v[i] : v[i + 4], keep = u4 ? v[i + 4] : v[i]; w4[i] = keep + __shfl_xor_sync(FULL, snd, 16); }
#pragma unroll
    for (int i = 0; i < 2; ++i) { const float snd = u3 ? w4[i] : w4[i + 2], keep = u3 ? w4[i + 2] : w4[i]; w2[i] = keep + __shfl_xor_sync(FULL, snd, 8); }
    const float snd = u2 ? w2[0] : w2[1], keep = u2 ? w2[1] : w2[0];
    float r = keep + __shfl_xor_sync(FULL, snd, 4);
    r += __shfl_xor_sync(FULL, r, 2);
    r += __shfl_xor_sync(FULL, r, 1);
    return r;
}

// ---- quarter-0 threads: step-size control, retire / refill, set-up of the next step (all in one phase between two CTA barriers) ----
__device__ __forceinline__ void atc_control(const AdjParams& P, AtcShared& S, int c, bool& queue_empty) {
    int col = S.col[c];
    int fl = 0;
    if (col >= 0) {
        fl = S.flag[c] & AFL_LAST;
        int status = S.status[c];
        if (S.bad[c]) { status = ST_NAN; S.mode[c] = 0; }              // non-finite lambda: its contributions were zeroed, the column is retired
        else if (S.mode[c]) {                                           // cancel pass done: retry with the reduced step
            S.mode[c] = 0;
            S.dt[c] = S.dt_retry[c];
        } else {
            const float ee = fast_sqrt(((S.part[0][c] + S.part[1][c]) + (S.part[2][c] + S.part[3][c])) * (1.f / NZ));      // fast controller arithmetic as in the forward kernels (nfc_tc.cuh)
            const float h = S.h[c];
            bool accept;
            float dtnew;
            if (P.fixed_dt > 0.f) { accept = true; dtnew = P.fixed_dt; }
            else if (!isfinite(ee)) { accept = false; dtnew = 0.2f * h; }
            else {
                pi_controller(ee, S.qold[c], h, accept, dtnew);
            }
            if (accept) {
                fl |= AFL_ACCEPT;
                S.t[c] = (fl & AFL_LAST) ? S.tstop[c] : S.t[c] - h;
                adj_accept_update(ee, h, S.dt[c], dtnew, P.fixed_dt <= 0.f, S.qold[c], S.dt[c]);
                S.nacc[c] += 1;
                S.rejrow[c] = 0;
                if (fl & AFL_LAST) {
                    if (S.si[c] >= 1) {
                        fl |= AFL_JUMP;
                        const int si = S.si[c];
                        S.jsi[c] = si;
                        S.si[c] = si - 1;
                        S.tstop[c] = P.saveat[si - 1 >= 0 ? si - 1 : 0];
                    } else fl |= AFL_DONE;
                }
            } else {
                S.nrej[c] += 1;
                S.mode[c] = 1;
                dtnew = adj_reject_dt(ee, h, dtnew, S.rejrow[c]);
                S.rejrow[c] += 1;
                S.dt_retry[c] = dtnew;
                if (dtnew < 1e-12f) status = ST_DT_UNDERFLOW;
            }
            if (S.nacc[c] + S.nrej[c] >= P.max_iters && !(fl & AFL_DONE) && status == ST_OK) status = ST_MAXITERS;
        }
        S.status[c] = status;
        if (status != ST_OK && S.mode[c] == 0) fl |= AFL_DONE;         // a failed column still runs its pending cancel pass first
        if (fl & AFL_DONE) {
            if (status != ST_OK && P.status) P.status[col] = status;
            if (P.adj_steps) P.adj_steps[col] = S.nacc[c] + S.nrej[c];
            atomicAdd(&P.counters[CTR_ADJ_ACCEPT], (unsigned long long)S.nacc[c]);
            atomicAdd(&P.counters[CTR_ADJ_REJECT], (unsigned long long)S.nrej[c]);
            col = -1;
        }
    }
    if (col < 0) {                          // empty slot: next column of the queue
        if (S.col[c] >= 0 || !queue_empty) fl |= AFL_NEWCOL;
        S.col[c] = -1; S.h[c] = 0.f; S.kk[c] = 0.f; S.mode[c] = 0; S.tlen[c] = 0; S.ipos[c] = 0; S.t[c] = 0.f; S.tstop[c] = 0.f;
        while (!queue_empty) {
            long long next = (long long)atomicAdd(&P.counters[P.queue_ctr], 1ULL);
            if (next >= P.B) { queue_empty = true; break; }
            if (P.order) next = P.order[next];
            const int len = P.tape_len[next];
            const int st = P.status ? P.status[next] : 0;
            if (st != ST_OK || len <= 0 || P.n_save < 2) continue;      // nothing to differentiate (or forward failed: reported in status[])
            const long long r0 = tape_row0(P, next);
            const float* last = P.tape + (r0 + (len - 1)) * TAPE_STRIDE;
            col = (int)next;
            S.row0[c] = r0;
            S.col[c] = col; S.t[c] = P.tf; S.tstop[c] = P.saveat[P.n_save - 2]; S.si[c] = P.n_save - 2;
            S.dt[c] = P.fixed_dt > 0.f ? P.fixed_dt : last[TAPE_H]; S.qold[c] = 1e-4f; S.nacc[c] = 0; S.nrej[c] = 0; S.rejrow[c] = 0;
            S.mode[c] = 0; S.ipos[c] = len - 1; S.tlen[c] = len; S.status[c] = ST_OK; S.dt_retry[c] = 0.f;
            S.kk[c] = P.use_kca ? P.colp[next * 3 + 2] * (float)NZ : 0.f;
            break;
        }
    }
    // next step: step length, tstop logic, tape interval of every stage (the stage times decrease monotonically: one walk down the tape)
    if (col >= 0) {
        float h = S.h[c];
        if (S.mode[c] == 0) {
            const float t = S.t[c], ts = S.tstop[c];
            h = fminf(S.dt[c], t - ts);
            fl &= ~AFL_LAST;
            if (t - h <= ts + 1e-6f * P.tf && (t - h) - ts <= 0.05f * fabsf(h)) { h = t - ts; fl |= AFL_LAST; } // snap to the save time only when that stretches the step by < 5 %: a wider window re-stretched a rejected step for ever (livelock at t - tstop = 4e-6, 2 of 10^6 columns)
            S.h[c] = h;
        }
        const float* base = P.tape + S.row0[c] * TAPE_STRIDE;
        int n = S.ipos[c];
        const int len = S.tlen[c];
        float tn = base[(long long)n * TAPE_STRIDE + TAPE_T], hn = base[(long long)n * TAPE_STRIDE + TAPE_H];
        const float t = S.t[c];
        int n0 = n;
#pragma unroll 1
        for (int s = 0; s < 7; ++s) {
            const float ts = (s == 6 && (fl & AFL_LAST)) ? S.tstop[c] : t - c_adj_c[s] * h;
            while (n > 0 && ts < tn) { --n; tn = base[(long long)n * TAPE_STRIDE + TAPE_T]; hn = base[(long long)n * TAPE_STRIDE + TAPE_H]; }
            while (n + 1 < len && ts > tn + hn) { ++n; tn = base[(long long)n * TAPE_STRIDE + TAPE_T]; hn = base[(long long)n * TAPE_STRIDE + TAPE_H]; }
            if (s == 0) n0 = n;
            S.ivp[s][c] = (short)n;
        }
        S.ipos[c] = n0;                     // a rejected step starts again from the same place
        S.any = 1;
    } else {
#pragma unroll
        for (int s = 0; s < 7; ++s) S.ivp[s][c] = 0;
    }
    S.flag[c] = fl;
    S.bad[c] = 0;
}

// Flush this step's weight-gradient accumulators (224 TMEM columns, lane = unit) into the CTA's gradient row with fire-and-forget RED.ADD
__device__ __forceinline__ void flush_acc(float* __restrict__ g, uint32_t tl, int c, int q, const AtcScales& sc) {
    using N0 = Net<128, 2, 0>;
#pragma unroll 1
    for (int gidx = q; gidx < 28; gidx += 4) {
        float v[8];
        tmem_ld8f(tl + 8 * gidx, v);
        wait_ld();
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int ci = 8 * gidx + e;        // accumulator column; TMEM lane c = unit j
            if (ci < 16) { if (ci == 0) atomicAdd(g + N0::T_HID + 128 * 128 + c, v[e]); }
            else if (ci < 144) atomicAdd(g + N0::T_HID + (ci - 16) * 128 + c, v[e] * sc.inv_sa1);
            else if (ci < 176) atomicAdd(g + N0::T_W0 + (ci - 144) * 128 + c, v[e] * sc.inv_sa0);
            else if (ci < 192) { if (ci == 176) atomicAdd(g + N0::T_B0 + c, v[e]); }
            else if (ci - 192 < 31) atomicAdd(g + N0::T_WL + c * 31 + (ci - 192), v[e] * sc.inv_sa2);
        }
    }
}

// u(t_s) of the CTA's 128 columns from the forward tape, loaded COOPERATIVELY and one stage ahead: warp w evaluates columns 8w..8w+7 with
// lane == level, so every load is one coalesced 128-B line (768 line requests per stage) and its latency hides under the MMAs of the
// stage before.  Loading with thread == column (each lane its own row: 6 144 sector requests per stage) saturated the L1 miss queue: 40 %
// of all stall samples.  Batch b = columns 8w + 4b .. 8w + 4b + 3.
__device__ __forceinline__ void coop_load(const AdjParams& P, AtcShared& S, unsigned char* smem, int s, int b, int warp, int lane) {
    float* ub = reinterpret_cast<float*>(smem + SO_UB);
    const float cs = c_adj_c[s];
    float e[4][5], tn[4], hn[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int cc = 8 * warp + 4 * b + j, colx = S.col[cc];
        const float* row = P.tape + ((colx >= 0 ? S.row0[cc] : 0LL) + S.ivp[s][cc]) * TAPE_STRIDE;
        tn[j] = __ldg(row + TAPE_T); hn[j] = __ldg(row + TAPE_H);
#pragma unroll
        for (int k = 0; k < 5; ++k) e[j][k] = __ldg(row + k * NZ + lane);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int cc = 8 * warp + 4 * b + j;
        const float hc = S.h[cc];
        const float ts = (s == 6 && (S.flag[cc] & AFL_LAST)) ? S.tstop[cc] : S.t[cc] - cs * hc;
        const float th = fminf(fmaxf(fast_div(ts - tn[j], hn[j]), 0.f), 1.f);
        const float uv = fmaf(hn[j], th * fmaf(th, fmaf(th, fmaf(th, e[j][4], e[j][3]), e[j][2]), e[j][1]), e[j][0]);
        ub[cc * 32 + ((((lane >> 2) ^ (cc & 7)) << 2) | (lane & 3))] = S.col[cc] >= 0 ? uv : 0.f;
    }
}

__global__ void __launch_bounds__(NTHREADS, 1) adjoint_tc_kernel(const unsigned char* __restrict__ gimg, const AdjParams P, const AtcScales sc) {
    extern __shared__ __align__(1024) unsigned char smem[];
    AtcShared& S = *reinterpret_cast<AtcShared*>(smem + SO_ST);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        mbar_init(&S.bar_w, 1);
        mbar_init(&S.bar_a, NCT);
        mbar_init(&S.bar_a1, NCT);
        mbar_init(&S.bar_d, 1);
        mbar_init(&S.bar_x, 1);
        S.stop = 0; S.any = 0;
    }
    __syncthreads();
    if (warp == NCT / 32) tmem_alloc(&S.tmem_base, TM_COLS);
    if (threadIdx.x == 0) {
        mbar_expect_tx(&S.bar_w, IMG16_BYTES);
        for (unsigned off = 0; off < (unsigned)IMG16_BYTES; off += 32768u) {
            const unsigned n = IMG16_BYTES - off < 32768u ? IMG16_BYTES - off : 32768u;
            bulk_g2s(smem + off, gimg + off, n, &S.bar_w);
        }
    }
    // staging buffers: zero (idle slots and the stage-7 pass never write them), constant-one rows of the bias gradients
    for (int i = threadIdx.x; i < (SO_UB + 16384 - SO_U) / 16; i += NTHREADS) reinterpret_cast<uint4*>(smem + SO_U)[i] = make_uint4(0u, 0u, 0u, 0u);
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * CH / 4; i += NTHREADS) reinterpret_cast<uint32_t*>(smem + SO_ONES)[i] = 0x3f803f80u;
    for (int i = threadIdx.x; i < TILE; i += NTHREADS) { S.col[i] = -1; S.bad[i] = 0; S.flag[i] = 0; S.mode[i] = 0; S.status[i] = ST_OK; }
    mbar_wait(&S.bar_w, 0);
    fence_async_smem();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t sbase = smem_u32(smem);
    const float* sb = reinterpret_cast<const float*>(smem + OFF16_B);      // bias1'[128] bias2'[128] b3[32] consts

    if (warp >= NCT / 32) {
        reg_dec<REGS_MMA>();
        if (warp == NCT / 32 && lane == 0) atc_mma_loop(&S, sbase);
        __syncwarp();
    } else {
        reg_inc<REGS_COL>();
        const int wq = warp & 3, q = warp >> 2, c = 32 * wq + lane;
        const uint32_t tl = S.tmem_base + ((uint32_t)(32 * wq) << 16);
        const float c0 = P.pref * (float)NZ;
        const float c1 = sb[288], c2 = sb[289], s_a0 = sb[291];
        {   // clear kappa_1 (stale stage vectors meet zero tableau entries and must be finite)
            const uint32_t z8[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
            tmem_st8(tl + TA_K1 + 8 * q, z8);
            wait_st();
        }
        float* const g = P.gacc + (size_t)blockIdx.x * P.n_params;
        int pd = 0, px = 0, col = -1;
        float lam[8], kr[5][8], uca[8], k7[8], lnew[8];
        float h = 0.f, sgn = 1.f, kk = 0.f, db3 = 0.f;
        uint32_t mask1 = 0u, mask2 = 0u;
        bool queue_empty = false;
#pragma unroll
        for (int i = 0; i < 8; ++i) { lam[i] = 0.f; uca[i] = 0.f; k7[i] = 0.f; lnew[i] = 0.f; kr[0][i] = 0.f; kr[1][i] = 0.f; kr[2][i] = 0.f; kr[3][i] = 0.f; kr[4][i] = 0.f; }
        fence_before_sync();
        cta_bar();
        while (true) {
            // ---------------- control phase (quarter-0 threads), then every thread picks up its column's new state ----------------
            if (q == 0) atc_control(P, S, c, queue_empty);
            cta_bar();
            {
                const int fl = S.flag[c];
                if (fl & AFL_ACCEPT) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) lam[i] = lnew[i];
                    if (fl & AFL_JUMP) {
                        const float* r = P.resid + ((long long)col * P.n_save + S.jsi[c]) * NZ + 8 * q;
                        const float4 a = *reinterpret_cast<const float4*>(r), b = *reinterpret_cast<const float4*>(r + 4);
                        lam[0] += a.x; lam[1] += a.y; lam[2] += a.z; lam[3] += a.w; lam[4] += b.x; lam[5] += b.y; lam[6] += b.z; lam[7] += b.w;
                    }
                }
                const bool nc = fl & AFL_NEWCOL;
                // tcgen05.st is warp-collective: kappa_1 of the whole warp is cleared when any lane changes column (all kappas are dead here)
                if (__any_sync(FULL, nc)) {
                    const uint32_t z8[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
                    tmem_st8(tl + TA_K1 + 8 * q, z8);
                    wait_st();
                }
                col = S.col[c];
                if (nc) {
                    if (col >= 0) {
                        const float* r = P.resid + ((long long)col * P.n_save + (P.n_save - 1)) * NZ + 8 * q;      // jump at tf
                        const float4 a = *reinterpret_cast<const float4*>(r), b = *reinterpret_cast<const float4*>(r + 4);
                        lam[0] = a.x; lam[1] = a.y; lam[2] = a.z; lam[3] = a.w; lam[4] = b.x; lam[5] = b.y; lam[6] = b.z; lam[7] = b.w;
                    } else {
#pragma unroll
                        for (int i = 0; i < 8; ++i) lam[i] = 0.f;
                    }
#pragma unroll
                    for (int i = 0; i < 8; ++i) { kr[0][i] = 0.f; kr[1][i] = 0.f; kr[2][i] = 0.f; kr[3][i] = 0.f; kr[4][i] = 0.f; }
                }
                h = S.h[c]; sgn = S.mode[c] ? -1.f : 1.f; kk = S.kk[c];
                // the backward solve walks down the tape: pull the two rows below this step's last interval into L2 now (a row is 6 lines;
                // the four threads of a column share them out), so that the next step's stage loads do not wait for HBM
                if (col >= 0) {
                    const int n6 = S.ivp[6][c];
                    const long long rb = S.row0[c];
                    const char* r0 = reinterpret_cast<const char*>(P.tape + (rb + (n6 > 0 ? n6 - 1 : 0)) * TAPE_STRIDE);
                    const char* r1 = reinterpret_cast<const char*>(P.tape + (rb + (n6 > 1 ? n6 - 2 : 0)) * TAPE_STRIDE);
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(r0 + 128 * q));
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(r1 + 128 * q));
                    if (q < 2) {
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(r0 + 128 * (q + 4)));
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(r1 + 128 * (q + 4)));
                    }
                }
            }
            const int any = S.any;
            if (any) { coop_load(P, S, smem, 0, 0, warp, lane); coop_load(P, S, smem, 0, 1, warp, lane); }
            cta_bar();
            if (threadIdx.x == 0) S.any = 0;          // re-armed for the next control phase (two barriers away)
            if (!any) break;

            // ------------------------------------------------ 7 stages ------------------------------------------------
#pragma unroll 1
            for (int s = 0; s < 7; ++s) {
                const bool dw = s < 6;
                const float w = (dw && col >= 0) ? sgn * h * c_adj_b[s] : 1.f;
                const float winv = 1.f / w;
                float scol = 1.f;           // per-column power-of-two normalisation of the back-propagated vector
                uint32_t ph3[4], pl3[4];
                // ---- E0: stage input lambda_s, u(t_s) from the tape, layer-1 operand ----
                {
                    float acc[8], kv[8];
                    tmem_ld8f(tl + TA_K1 + 8 * q, kv);
                    const float a0 = c_adj_a[s][0], a1 = c_adj_a[s][1], a2 = c_adj_a[s][2], a3 = c_adj_a[s][3], a4 = c_adj_a[s][4], a5 = c_adj_a[s][5];
                    float u[8], lw[8];
                    {   // u(t_s): put into the u buffer one stage ahead by coop_load (below)
                        const float* ub = reinterpret_cast<const float*>(smem + SO_UB) + c * 32;
                        const float4 ua = *reinterpret_cast<const float4*>(ub + (((2 * q) ^ (c & 7)) << 2)), uc = *reinterpret_cast<const float4*>(ub + (((2 * q + 1) ^ (c & 7)) << 2));
                        u[0] = ua.x; u[1] = ua.y; u[2] = ua.z; u[3] = ua.w; u[4] = uc.x; u[5] = uc.y; u[6] = uc.z; u[7] = uc.w;
                    }
                    wait_ld();
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        float a = a0 * kv[i];
                        a = fmaf(a1, kr[0][i], a); a = fmaf(a2, kr[1][i], a); a = fmaf(a3, kr[2][i], a); a = fmaf(a4, kr[3][i], a); a = fmaf(a5, kr[4][i], a);
                        acc[i] = fmaf(h, a, lam[i]);            // lambda_s
                    }
                    float m = 0.f, nanacc = 0.f;
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        lw[i] = w * acc[i];
                        m = fmaxf(m, fabsf(lw[i]));
                        nanacc = fmaf(lw[i], 0.f, nanacc);
                    }
                    S.xu[q][0][c] = u[0]; S.xu[q][1][c] = u[7];
                    S.xl[q][0][c] = lw[0]; S.xl[q][1][c] = lw[7];
                    S.xm[q][c] = nanacc == 0.f ? m : __int_as_float(0x7fc00000);
                    // layer-1 A operand (the network input saturates at the operand bound like in the forward solve)
                    uint32_t ph[4], pl[4], pb[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float v0 = fminf(fmaxf(u[2 * i], -SplitFP16x2::INPUT_BOUND), SplitFP16x2::INPUT_BOUND) * s_a0;
                        const float v1 = fminf(fmaxf(u[2 * i + 1], -SplitFP16x2::INPUT_BOUND), SplitFP16x2::INPUT_BOUND) * s_a0;
                        split2h(v0, v1, ph[i], pl[i]);
                        pb[i] = pack_bf16(v0, v1);
                    }
                    tmem_st4(tl + TA_A + 4 * q, ph);
                    tmem_st4(tl + TA_A + 64 + 4 * q, pl);
                    if (s > 0) { mbar_wait(&S.bar_x, px); px ^= 1; }      // dW1 of the previous stage has read the u staging buffer
                    if (dw) *reinterpret_cast<uint4*>(smem + SO_U + q * CH + c * 16) = make_uint4(pb[0], pb[1], pb[2], pb[3]);
                    wait_st();
                    fence_before_sync();
                    fence_async_smem();
                    mbar_arrive(&S.bar_a);
                    mbar_arrive(&S.bar_a1);
                    // keep w*lambda_s and u for the face quantities (computed after the exchange, below)
#pragma unroll
                    for (int i = 0; i < 8; ++i) { uca[i] = u[i]; lnew[i] = lw[i]; }
                }
                // ---- E1: z1 -> a1 (operand of layer 2, staged for dW2, relu mask); face quantities: output-layer delta, K_CA Jacobian ----
                mbar_wait(&S.bar_d, pd); pd ^= 1;
                fence_after_sync();
                {
                    // the exchange buffers were written before the barrier traffic of the layer-1 MMAs
                    const float ulo = q > 0 ? S.xu[q > 0 ? q - 1 : 0][1][c] : 0.f, uhi = q < 3 ? S.xu[q < 3 ? q + 1 : 3][0][c] : 0.f;
                    const float llo = q > 0 ? S.xl[q > 0 ? q - 1 : 0][1][c] : 0.f, lhi = q < 3 ? S.xl[q < 3 ? q + 1 : 3][0][c] : 0.f;
                    const float m0 = S.xm[0][c], m1 = S.xm[1][c], m2 = S.xm[2][c], m3 = S.xm[3][c];
                    const float M = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
                    const bool isbad = !(m0 == m0 && m1 == m1 && m2 == m2 && m3 == m3) || !(M < 1e30f);
                    {
                        // |dL/dF_k| = c0 |lw_k - lw_{k-1}| <= 2 c0 M  ->  scol = 2^e with 2 c0 M scol in [1/2, 1)
                        const float bound = 2.f * fabsf(c0) * M;
                        int E = (__float_as_int(bound) >> 23) & 0xff;
                        E = E < 66 ? 66 : E;                                  // tiny lambda: cap the scale at 2^60
                        scol = __int_as_float((253 - E) << 23);
                        if (isbad) scol = 1.f;
                    }
                    if (q == 0 && isbad) S.bad[c] = 1;
                    float Fb[9], sca[9], d3[8];
#pragma unroll
                    for (int f = 0; f <= 8; ++f) {
                        const int k = 8 * q + f;
                        const float lk = f < 8 ? lnew[f < 8 ? f : 0] : lhi, lkm = f > 0 ? lnew[f > 0 ? f - 1 : 0] : llo;
                        const float uk = f < 8 ? uca[f < 8 ? f : 0] : uhi, ukm = f > 0 ? uca[f > 0 ? f - 1 : 0] : ulo;
                        float F = c0 * (lk - lkm);
                        if (k == 0 || k == NZ || isbad) F = 0.f;
                        Fb[f] = F;
                        sca[f] = (kk * (uk - ukm) < 0.f) ? kk * F : 0.f;
                    }
#pragma unroll
                    for (int i = 0; i < 8; ++i) { uca[i] = sca[i + 1] - sca[i]; d3[i] = Fb[i + 1]; }
                    if (dw) {
                        db3 += warp_sum8(d3, lane);
                        *reinterpret_cast<uint4*>(smem + SO_D3 + q * CH + c * 16) =
                            make_uint4(pack_bf16(d3[0], d3[1]), pack_bf16(d3[2], d3[3]), pack_bf16(d3[4], d3[5]), pack_bf16(d3[6], d3[7]));
                    }
                    const float s3 = scol * sc.s_d3;
#pragma unroll
                    for (int i = 0; i < 4; ++i) split2h(d3[2 * i] * s3, d3[2 * i + 1] * s3, ph3[i], pl3[i]);
                }
                mask1 = 0u;
                {
                    // the whole D slice first: the layer-2 MMAs of the even k-steps overwrite D as soon as the first halves are published
                    uint32_t r[32];
                    tmem_ld16(tl + TA_D + 32 * q, *reinterpret_cast<uint32_t(*)[16]>(&r[0]));
                    tmem_ld16(tl + TA_D + 32 * q + 16, *reinterpret_cast<uint32_t(*)[16]>(&r[16]));
                    wait_ld();
#pragma unroll
                    for (int ch = 0; ch < 2; ++ch) {
                        uint32_t ph[8], pl[8], pb[8];
                        const float* b = sb + 32 * q + 16 * ch;
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const float2 bb = *reinterpret_cast<const float2*>(b + 2 * i);
                            const float v0 = fmaf(__uint_as_float(r[16 * ch + 2 * i]), c1, bb.x), v1 = fmaf(__uint_as_float(r[16 * ch + 2 * i + 1]), c1, bb.y);
                            mask_push(mask1, v0); mask_push(mask1, v1);
                            split2h_relu(v0, v1, ph[i], pl[i]);
                            pb[i] = pack_bf16_relu(v0, v1);
                        }
                        tmem_st8(tl + TA_A + 16 * q + 8 * ch, ph);
                        tmem_st8(tl + TA_A + 64 + 16 * q + 8 * ch, pl);
                        if (dw) {
                            *reinterpret_cast<uint4*>(smem + SO_A1 + (4 * q + 2 * ch) * CH + c * 16) = make_uint4(pb[0], pb[1], pb[2], pb[3]);
                            *reinterpret_cast<uint4*>(smem + SO_A1 + (4 * q + 2 * ch + 1) * CH + c * 16) = make_uint4(pb[4], pb[5], pb[6], pb[7]);
                        }
                        wait_st();
                        fence_before_sync();
                        if (ch == 0) mbar_arrive(&S.bar_a);
                        else { fence_async_smem(); mbar_arrive(&S.bar_a1); }
                    }
                }
                mask1 = ~mask1;
                if (s < 6) coop_load(P, S, smem, s + 1, 0, warp, lane);      // under the layer-2 MMAs
                else flush_acc(g, tl, c, q, sc);                             // every contraction of this step has completed (bar_x at E0)
                // ---- E2: z2 -> a2 (staged for dW3^T, relu mask); output-layer delta -> A operand ----
                mbar_wait(&S.bar_d, pd); pd ^= 1;
                fence_after_sync();
                mask2 = 0u;
                {
                    uint32_t r[32];
                    tmem_ld16(tl + TA_D + 32 * q, *reinterpret_cast<uint32_t(*)[16]>(&r[0]));
                    tmem_ld16(tl + TA_D + 32 * q + 16, *reinterpret_cast<uint32_t(*)[16]>(&r[16]));
                    wait_ld();
#pragma unroll
                    for (int ch = 0; ch < 2; ++ch) {
                        uint32_t pb[8];
                        const float* b = sb + 128 + 32 * q + 16 * ch;
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const float2 bb = *reinterpret_cast<const float2*>(b + 2 * i);
                            const float v0 = fmaf(__uint_as_float(r[16 * ch + 2 * i]), c2, bb.x), v1 = fmaf(__uint_as_float(r[16 * ch + 2 * i + 1]), c2, bb.y);
                            mask_push(mask2, v0); mask_push(mask2, v1);
                            pb[i] = pack_bf16_relu(v0, v1);
                        }
                        if (dw) {
                            *reinterpret_cast<uint4*>(smem + SO_X + (4 * q + 2 * ch) * CH + c * 16) = make_uint4(pb[0], pb[1], pb[2], pb[3]);
                            *reinterpret_cast<uint4*>(smem + SO_X + (4 * q + 2 * ch + 1) * CH + c * 16) = make_uint4(pb[4], pb[5], pb[6], pb[7]);
                        }
                    }
                }
                mask2 = ~mask2;
                tmem_st4(tl + TA_A + 4 * q, ph3);
                tmem_st4(tl + TA_A + 64 + 4 * q, pl3);
                wait_st();
                fence_before_sync();
                fence_async_smem();
                mbar_arrive(&S.bar_a);
                mbar_arrive(&S.bar_a1);
                // ---- E4 / E5: g2 -> d2 = g2 . relu'(z2), g1 -> d1 = g1 . relu'(z1): next operand + staged (true scale) for dW2 / dW1 ----
#pragma unroll 1
                for (int l = 0; l < 2; ++l) {
                    mbar_wait(&S.bar_d, pd); pd ^= 1;
                    fence_after_sync();
                    const float cg = l == 0 ? sc.c_g2 : sc.c_g1;
                    const float cst = (l == 0 ? sc.inv_sd2 : sc.inv_sd1) / scol;      // a power of two: exact in bf16
                    const uint32_t cst2 = pack_bf16(cst, cst);
                    const uint32_t mk = l == 0 ? mask2 : mask1;
                    uint32_t pb[16];
                    {
                        uint32_t r[32];
                        tmem_ld16(tl + TA_D + 32 * q, *reinterpret_cast<uint32_t(*)[16]>(&r[0]));
                        tmem_ld16(tl + TA_D + 32 * q + 16, *reinterpret_cast<uint32_t(*)[16]>(&r[16]));
                        wait_ld();
#pragma unroll
                        for (int ch = 0; ch < 2; ++ch) {
                            uint32_t ph[8], pl[8];
#pragma unroll
                            for (int i = 0; i < 8; ++i) {
                                const float v0 = ((mk >> (31 - (16 * ch + 2 * i))) & 1u) ? __uint_as_float(r[16 * ch + 2 * i]) * cg : 0.f;
                                const float v1 = ((mk >> (31 - (16 * ch + 2 * i + 1))) & 1u) ? __uint_as_float(r[16 * ch + 2 * i + 1]) * cg : 0.f;
                                split2h(v0, v1, ph[i], pl[i]);
                                pb[8 * ch + i] = mul_bf16x2(pack_bf16(v0, v1), cst2);
                            }
                            tmem_st8(tl + TA_A + 16 * q + 8 * ch, ph);
                            tmem_st8(tl + TA_A + 64 + 16 * q + 8 * ch, pl);
                            if (ch == 0) { wait_st(); fence_before_sync(); mbar_arrive(&S.bar_a); }
                        }
                    }
                    if (dw) {
                        mbar_wait(&S.bar_x, px); px ^= 1;          // the previous contraction (issued behind this layer's MMAs) has read the X staging buffer
#pragma unroll
                        for (int g4 = 0; g4 < 4; ++g4)
                            *reinterpret_cast<uint4*>(smem + SO_X + (4 * q + g4) * CH + c * 16) = make_uint4(pb[4 * g4], pb[4 * g4 + 1], pb[4 * g4 + 2], pb[4 * g4 + 3]);
                    }
                    wait_st();
                    fence_before_sync();
                    fence_async_smem();
                    mbar_arrive(&S.bar_a1);
                    if (l == 0 && s < 6) coop_load(P, S, smem, s + 1, 1, warp, lane);      // under the g1 MMAs
                }
                // ---- E6: kappa_s = (W1^T d1 + K_CA part) / w ----
                mbar_wait(&S.bar_d, pd); pd ^= 1;
                fence_after_sync();
                {
                    float a[8], b[8], kap[8];
                    tmem_ld8f(tl + TA_D + 8 * q, a);
                    tmem_ld8f(tl + TA_D + 32 + 8 * q, b);
                    wait_ld();
                    const float cu = sc.inv_u / scol;
#pragma unroll
                    for (int i = 0; i < 8; ++i) kap[i] = fmaf(a[i] + b[i], cu, uca[i]) * winv;
                    if (s == 0) { tmem_st8f(tl + TA_K1 + 8 * q, kap); wait_st(); }
                    else {
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            if (s == 1) kr[0][i] = kap[i];
                            else if (s == 2) kr[1][i] = kap[i];
                            else if (s == 3) kr[2][i] = kap[i];
                            else if (s == 4) kr[3][i] = kap[i];
                            else if (s == 5) kr[4][i] = kap[i];
                            else k7[i] = kap[i];
                        }
                    }
                }
                fence_before_sync();
            }
            // ---- step end: lambda_new, error estimate on lambda ----
            {
                float kv[8];
                tmem_ld8f(tl + TA_K1 + 8 * q, kv);
                wait_ld();
                float part = 0.f;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    float b = NFC_B1 * kv[i];
                    b = fmaf(NFC_B2, kr[0][i], b); b = fmaf(NFC_B3, kr[1][i], b); b = fmaf(NFC_B4, kr[2][i], b); b = fmaf(NFC_B5, kr[3][i], b); b = fmaf(NFC_B6, kr[4][i], b);
                    lnew[i] = fmaf(h, b, lam[i]);
                    float e = NFC_BT1 * kv[i];
                    e = fmaf(NFC_BT2, kr[0][i], e); e = fmaf(NFC_BT3, kr[1][i], e); e = fmaf(NFC_BT4, kr[2][i], e); e = fmaf(NFC_BT5, kr[3][i], e);
                    e = fmaf(NFC_BT6, kr[4][i], e); e = fmaf(NFC_BT7, k7[i], e);
                    e *= h;
                    const float r = fast_div(e, fmaf(fmaxf(fabsf(lam[i]), fabsf(lnew[i])), P.reltol, P.abstol));
                    part = fmaf(r, r, part);
                }
                S.part[q][c] = part;
            }
            cta_bar();
        }
        // ---- flush: bias gradient of the output layer (registers), then the persistent accumulators ----
        using N0 = Net<128, 2, 0>;
        if ((lane & 3) == 0) {
            const int n = 8 * q + ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);
            if (n < 31) atomicAdd(g + N0::T_BL + n, db3);
        }
        // release the MMA warp
        cta_bar();
        if (threadIdx.x == 0) S.stop = 1;
        cta_bar();
        mbar_arrive(&S.bar_a);
    }
    fence_before_sync();
    __syncthreads();
    if (warp == NCT / 32) tmem_dealloc(S.tmem_base, TM_COLS);
}

}  // namespace atc
}  // namespace nfc
