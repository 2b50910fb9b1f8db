// nfc_tc.cuh -- tensor-core (tcgen05 / TMEM) evaluation of the Chain for a tile of 128 ocean columns.
//
// One CTA = one tile of 128 columns = the M dimension of a tcgen05.mma (M = 128 TMEM lanes, lane == ocean column).
// Dense layer l:  D[128 x N_l] (fp32, TMEM) = A[128 x K_l] * W_l^T   with
//   * A (activations of the 128 columns) living in TENSOR MEMORY as the A operand (TS form of tcgen05.mma): the
//     epilogue threads write it with tcgen05.st, so activations never touch shared memory;
//   * W_l in shared memory as the K-major, no-swizzle UMMA B operand (8 x 16-byte core matrices), staged once per CTA
//     by the bulk-copy engine;
//   * FP32-faithful arithmetic from BF16 tensor cores: every fp32 value is split into three bf16 terms
//     x = x1 + x2 + x3 (8+8+8 significand bits) and the product keeps the six terms down to 2^-16:
//     a1b1 + a1b2 + a2b1 + a2b2 + a1b3 + a3b1  (dropped terms <= 2^-24 relative) -- six kind::f16 MMAs per k-step,
//     accumulated in fp32 in TMEM.  BF16 keeps fp32's exponent range, so no scaling is needed (weights x1e-5 are fine).
// Thread roles: warps 0..15 = 512 "column" threads, FOUR per ocean column: warps w, w+4, w+8, w+12 own TMEM lanes
// 32 (w%4)..+31 and quarter q = w/4 owns levels [8q, 8q+8) and hidden units [32q, 32q+32) -- 4 column warps per SM
// sub-partition hide the epilogue's latencies.  Warps 16..19 = MMA warpgroup (warp 16, one elected lane, issues; the
// group hands its registers to the column threads with setmaxnreg).  The Runge-Kutta stage vectors k2..k7 of the solve
// kernel live in the 192 tensor-memory columns the GEMMs do not use, so the ODE state costs 16 registers per thread.
// Only the default Chain Dense(32,128,relu) -> Dense(128,128,relu) -> Dense(128,31) is instantiated here.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "nfc_device.cuh"

namespace nfc {
namespace tc {

constexpr int TILE = 128;                 // ocean columns per CTA
constexpr int H = 128;                    // hidden width
constexpr int NCT = 512;                  // column threads (4 per ocean column)
constexpr int NTHREADS = NCT + 128;       // + the MMA warpgroup (only its first warp works; a full warpgroup is needed for setmaxnreg)
constexpr int REGS_MMA = 32, REGS_COL = 112;   // launch: 96/thread; (96 - 32) * 128 freed == (112 - 96) * 512 claimed
// tensor-memory columns (32-bit cells per lane)
constexpr int TM_D = 0;                   // accumulator, up to 128 columns
constexpr int TM_A = 128;                 // A operand: 3 splits x 64 columns (K = 128 bf16 packed two per cell)
constexpr int TM_K = 320;                 // k2..k7 of the ODE solve: 6 x 32 columns (level = column offset)
constexpr int TM_COLS = 512;
// shared-memory image (bytes): three bf16 splits of each weight matrix in UMMA K-major no-swizzle layout, then biases
constexpr int W1_BYTES = 128 * 32 * 2;    // N = 128, K = 32
constexpr int W2_BYTES = 128 * 128 * 2;   // N = 128, K = 128
constexpr int W3_BYTES = 32 * 128 * 2;    // N = 32 (31 + zero row), K = 128
constexpr int OFF_W1 = 0;
constexpr int OFF_W2 = OFF_W1 + 3 * W1_BYTES;
constexpr int OFF_W3 = OFF_W2 + 3 * W2_BYTES;
constexpr int OFF_B = OFF_W3 + 3 * W3_BYTES;           // b1[128] b2[128] b3[32] fp32
constexpr int IMG_BYTES = OFF_B + (128 + 128 + 32) * 4;   // 148 608 B

// byte offset of element (n, k) inside one split image of an [N][K] matrix: K chunks of 8 elements are LBO apart,
// groups of 8 rows are SBO = 128 B apart, a core matrix is 8 rows x 16 B
__host__ __device__ constexpr int umma_off(int n, int k, int N) { return (k >> 3) * (N * 16) + (n >> 3) * 128 + (n & 7) * 16 + (k & 7) * 2; }

// ---- PTX wrappers ----------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
template <int R> __device__ __forceinline__ void reg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(R)); }
template <int R> __device__ __forceinline__ void reg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(R)); }
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// one arrival per WARP (the barriers of the A operand count the 16 column warps): 512 per-thread arrivals on one shared-memory word
// serialise and sat on the critical path of every hand-off to the MMA warp.  Each lane has completed its tcgen05.st (wait::st) and issued
// tcgen05.fence::before_thread_sync; the warp barrier orders the lanes before lane 0 arrives for all of them.
constexpr int NCW = 16;                   // column warps
__device__ __forceinline__ void warp_arrive(uint64_t* bar) {
    __syncwarp();
    if ((threadIdx.x & 31) == 0) mbar_arrive(bar);
}
// D[tmem] (+)= A[tmem] * B[smem desc]   (kind::f16, bf16 inputs, fp32 accumulate)
__device__ __forceinline__ void mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
        "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// same, descriptor given as (lo, hi) words: per k-step only the 14-bit address field in the low word moves
template <bool ACC>
__device__ __forceinline__ void mma_ts2(uint32_t d_tmem, uint32_t a_tmem, uint32_t b_lo, uint32_t b_hi, uint32_t idesc) {
    if (ACC)
        asm volatile("{\n\t.reg .b64 bd;\n\t.reg .pred p;\n\tsetp.eq.u32 p, 1, 1;\n\tmov.b64 bd, {%2, %3};\n\t"
                     "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], bd, %4, p;\n\t}" ::"r"(d_tmem), "r"(a_tmem), "r"(b_lo), "r"(b_hi), "r"(idesc) : "memory");
    else
        asm volatile("{\n\t.reg .b64 bd;\n\t.reg .pred p;\n\tsetp.ne.u32 p, 1, 1;\n\tmov.b64 bd, {%2, %3};\n\t"
                     "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], bd, %4, p;\n\t}" ::"r"(d_tmem), "r"(a_tmem), "r"(b_lo), "r"(b_hi), "r"(idesc) : "memory");
}
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
          "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]),
                 "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                 : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr)
                 : "memory");
}
__device__ __forceinline__ void tmem_st4(uint32_t taddr, const uint32_t (&r)[4]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]) : "memory");
}
__device__ __forceinline__ void tmem_ld8f(uint32_t taddr, float (&f)[8]) {
    uint32_t r[8];
    tmem_ld8(taddr, r);
#pragma unroll
    for (int i = 0; i < 8; ++i) f[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_st8f(uint32_t taddr, const float (&f)[8]) {
    uint32_t r[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) r[i] = __float_as_uint(f[i]);
    tmem_st8(taddr, r);
}
__device__ __forceinline__ void tmem_ld1(uint32_t taddr, uint32_t& r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(r) : "r"(taddr) : "memory");
}

__host__ __device__ constexpr uint32_t make_idesc(int M, int N) {   // kind::f16: A = B = BF16 (1), C = F32 (1), both K-major
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ uint64_t make_bdesc(uint32_t smem_addr, int N) {   // K-major, SWIZZLE_NONE: LBO = N*16 B, SBO = 128 B
    return (uint64_t)((smem_addr >> 4) & 0x3FFF) | ((uint64_t)(((uint32_t)N * 16) >> 4) << 16) | ((uint64_t)(128 >> 4) << 32) | (1ull << 46);
}

// ---- packed fp32 (sm_100: FFMA2 / FADD2 / FMUL2, two IEEE fp32 operations per instruction, results identical to the scalar forms).
// The column threads are bound by instruction issue and dependent-issue latency, not by the FP32 pipe, so halving the number of FMA / ADD
// instructions of the 8-level vectors is a direct gain.  Pairs are (even, odd) elements; ptxas keeps them in adjacent registers.
__device__ __forceinline__ unsigned long long pk2(float lo, float hi) { unsigned long long r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void upk2(unsigned long long v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
// (d0, d1) = (a0, a1) * s + (c0, c1)
__device__ __forceinline__ void fma2s(float& d0, float& d1, float a0, float a1, float s, float c0, float c1) {
    unsigned long long r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(pk2(a0, a1)), "l"(pk2(s, s)), "l"(pk2(c0, c1)));
    upk2(r, d0, d1);
}
// (d0, d1) = (a0, a1) * (b0, b1) + (c0, c1)
__device__ __forceinline__ void fma2(float& d0, float& d1, float a0, float a1, float b0, float b1, float c0, float c1) {
    unsigned long long r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(pk2(a0, a1)), "l"(pk2(b0, b1)), "l"(pk2(c0, c1)));
    upk2(r, d0, d1);
}
__device__ __forceinline__ void add2(float& d0, float& d1, float a0, float a1, float b0, float b1) {
    unsigned long long r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(pk2(a0, a1)), "l"(pk2(b0, b1)));
    upk2(r, d0, d1);
}
__device__ __forceinline__ void sub2(float& d0, float& d1, float a0, float a1, float b0, float b1) {
    unsigned long long r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(pk2(a0, a1)), "l"(pk2(b0, b1)));
    upk2(r, d0, d1);
}
__device__ __forceinline__ void mul2s(float& d0, float& d1, float a0, float a1, float s) {
    unsigned long long r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(pk2(a0, a1)), "l"(pk2(s, s)));
    upk2(r, d0, d1);
}
// acc[0..8) = coef * k[0..8) + acc[0..8)   /   acc = coef * k
__device__ __forceinline__ void axpy8(float (&acc)[8], float coef, const float (&k)[8]) {
#pragma unroll
    for (int i = 0; i < 8; i += 2) fma2s(acc[i], acc[i + 1], k[i], k[i + 1], coef, acc[i], acc[i + 1]);
}
__device__ __forceinline__ void scal8(float (&acc)[8], float coef, const float (&k)[8]) {
#pragma unroll
    for (int i = 0; i < 8; i += 2) mul2s(acc[i], acc[i + 1], k[i], k[i + 1], coef);
}

// x = x1 + x2 + x3 in bf16; the pair (even, odd) is packed as (low, high) half of one 32-bit cell
__device__ __forceinline__ void split3(float v0, float v1, uint32_t& p1, uint32_t& p2, uint32_t& p3) {
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(p1) : "f"(v1), "f"(v0));
    const float r0 = v0 - __uint_as_float(p1 << 16), r1 = v1 - __uint_as_float(p1 & 0xffff0000u);
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(p2) : "f"(r1), "f"(r0));
    const float s0 = r0 - __uint_as_float(p2 << 16), s1 = r1 - __uint_as_float(p2 & 0xffff0000u);
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(p3) : "f"(s1), "f"(s0));
}

// ---- weight image: theta (Flux layout) -> three bf16 splits per layer in UMMA layout + fp32 biases ----------------------
static __global__ void pack_weights_tc_kernel(const float* __restrict__ theta, unsigned char* __restrict__ img) {
    using N = Net<128, 2, 0>;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int n1 = 128 * 32, n2 = 128 * 128, n3 = 32 * 128;
    float w = 0.f;
    int off = 0, bytes = 0;
    if (i < n1) { const int n = i % 128, k = i / 128; w = theta[N::T_W0 + k * 128 + n]; off = OFF_W1 + umma_off(n, k, 128); bytes = W1_BYTES; }
    else if (i < n1 + n2) { const int q = i - n1, n = q % 128, k = q / 128; w = theta[N::T_HID + k * 128 + n]; off = OFF_W2 + umma_off(n, k, 128); bytes = W2_BYTES; }
    else if (i < n1 + n2 + n3) {
        const int q = i - n1 - n2, n = q % 32, k = q / 32;
        w = n < 31 ? theta[N::T_WL + k * 31 + n] : 0.f;
        // output layer: the three splits are stacked along N (rows 32 s + n of ONE [96][128] operand), see issue_layer<2>
        const __nv_bfloat16 c1 = __float2bfloat16_rn(w);
        const float q1 = w - __bfloat162float(c1);
        const __nv_bfloat16 c2 = __float2bfloat16_rn(q1);
        const __nv_bfloat16 c3 = __float2bfloat16_rn(q1 - __bfloat162float(c2));
        *reinterpret_cast<__nv_bfloat16*>(img + OFF_W3 + umma_off(n, k, 96)) = c1;
        *reinterpret_cast<__nv_bfloat16*>(img + OFF_W3 + umma_off(32 + n, k, 96)) = c2;
        *reinterpret_cast<__nv_bfloat16*>(img + OFF_W3 + umma_off(64 + n, k, 96)) = c3;
        return;
    } else if (i < n1 + n2 + n3 + 288) {
        const int q = i - n1 - n2 - n3;
        float b = 0.f;
        if (q < 128) b = theta[N::T_B0 + q];
        else if (q < 256) b = theta[N::T_HID + 128 * 128 + (q - 128)];
        else if (q - 256 < 31) b = theta[N::T_BL + (q - 256)];
        reinterpret_cast<float*>(img + OFF_B)[q] = b;
        return;
    } else return;
    const __nv_bfloat16 b1 = __float2bfloat16_rn(w);
    const float r1 = w - __bfloat162float(b1);
    const __nv_bfloat16 b2 = __float2bfloat16_rn(r1);
    const __nv_bfloat16 b3 = __float2bfloat16_rn(r1 - __bfloat162float(b2));
    *reinterpret_cast<__nv_bfloat16*>(img + off) = b1;
    *reinterpret_cast<__nv_bfloat16*>(img + off + bytes) = b2;
    *reinterpret_cast<__nv_bfloat16*>(img + off + 2 * bytes) = b3;
}

// ---- shared state of one CTA ------------------------------------------------------------------------------------------
struct TcShared {
    uint64_t bar_w;        // weight image landed
    uint64_t bar_a;        // A operand, sub-chunk 0 written by all column threads: the whole layer-1 operand / k-steps 0, 1 of a later layer
    uint64_t bar_q[3];     // sub-chunks 1..3 of a hidden or output layer's operand (k-steps 2s, 2s+1)
    uint64_t bar_d;        // accumulator ready (tcgen05.commit)
    uint32_t tmem_base;
    int stop;
    int nprod;             // split products issued per k-step (6 = FP32-faithful; fewer only for timing experiments)
    float xch[2][TILE][4][2]; // first / last level of every quarter, exchanged between the four threads of a column (double buffered)
    float part[2][TILE][4];   // partial error norms (double buffered)
    float rng[2][TILE][4];    // max |u_new| of each quarter (range check of accepted states)
    float hal[2][TILE][4][3]; // first three levels of every quarter: the halo of the Conv front-end (conv-2 / conv-4 Chains)
    int newcol[TILE];         // refill hand-off between the four threads of a column
    int any;
};

// MMA issuer: the split products of one Dense layer, fully unrolled: per MMA one 32-bit add on the descriptor's
// address field (k-step stride 2*LBO), one add on the TMEM column, the tcgen05.mma itself.  (A first version built every
// descriptor in a runtime loop and was issue-bound at ~98 cycles per MMA against 64 / 16 cycles of tensor-pipe time.)
// SUB selects the k-steps 2 SUB, 2 SUB + 1, whose A cells the column threads publish as their sub-chunk SUB (layer 0: K = 32, one
// sub-chunk); the very first MMA of a layer overwrites D
template <int LAYER, int SUB>
__device__ __forceinline__ void issue_layer(uint32_t tmem_base, uint32_t smem_img, int nprod) {
    constexpr int Nn = LAYER == 2 ? 96 : 128;
    constexpr uint32_t woff = LAYER == 0 ? OFF_W1 : (LAYER == 1 ? OFF_W2 : OFF_W3);
    constexpr uint32_t wbytes = LAYER == 0 ? W1_BYTES : (LAYER == 1 ? W2_BYTES : W3_BYTES);
    constexpr uint32_t idesc = make_idesc(128, Nn);
    constexpr uint32_t kstep = (2u * Nn * 16u) >> 4;            // descriptor address-field increment per k-step (16 elements)
    const uint32_t d = tmem_base + TM_D;
    const uint64_t desc0 = make_bdesc(smem_img + woff, Nn);
    const uint32_t lo0 = (uint32_t)desc0, hi0 = (uint32_t)(desc0 >> 32);
    const uint32_t a_base = tmem_base + TM_A;
    // (a split, b split): smallest terms first so the big a1 b1 term lands on an already formed tail
    constexpr int sa[6] = {2, 0, 1, 1, 0, 0};
    constexpr int sb[6] = {0, 2, 1, 0, 1, 0};
    if (LAYER == 2) {
        // Output layer (31 neurons): a TS-form MMA costs ~61 cycles whatever N is (the A operand read from tensor memory, 4 KB per
        // k-step at 64 B/cycle, is the limit), so the three weight splits are stacked along N: one N = 96 MMA per A split,
        // D[:, 0:32] + D[:, 32:64] + D[:, 64:96] = A (B1 + B2 + B3) with all nine product terms, 24 MMAs instead of 48.
#pragma unroll
        for (int sA = 2; sA >= 0; --sA) {
#pragma unroll
            for (int j = 2 * SUB; j < 2 * SUB + 2; ++j) {
                if (sA == 2 && j == 0) mma_ts2<false>(d, a_base + sA * 64 + 8 * j, lo0 + kstep * j, hi0, idesc);
                else mma_ts2<true>(d, a_base + sA * 64 + 8 * j, lo0 + kstep * j, hi0, idesc);
            }
        }
    } else if (nprod >= 6) {
#pragma unroll
        for (int p = 0; p < 6; ++p) {
#pragma unroll
            for (int j = 2 * SUB; j < 2 * SUB + 2; ++j) {
                const uint32_t blo = lo0 + ((sb[p] * wbytes) >> 4) + kstep * j;      // compile-time offset, no carry out of the 14-bit field
                if (p == 0 && j == 0) mma_ts2<false>(d, a_base + sa[p] * 64 + 8 * j, blo, hi0, idesc);
                else mma_ts2<true>(d, a_base + sa[p] * 64 + 8 * j, blo, hi0, idesc);
            }
        }
    } else {   // timing experiments only
        uint32_t acc = SUB != 0;
        for (int p = 6 - nprod; p < 6; ++p)
            for (int j = 2 * SUB; j < 2 * SUB + 2; ++j) { mma_ts(d, a_base + sa[p] * 64 + 8 * j, desc0 + (uint64_t)(((sb[p] * wbytes) >> 4) + kstep * j), idesc, acc); acc = 1; }
    }
}

// hooks of a Chain evaluation: work the column threads do while the tensor pipe runs layer 2 (H1) / the output layer (H2)
struct NoHook { __device__ __forceinline__ void operator()() const {} };

// Column-thread side of one Chain evaluation.  y[8]: this thread's 8 levels of the stage input (level 8 q + i).
// Returns nn[i] = NN output n = 8 q + i - 1  (i = 0..8; nn[0] of q = 0 and nn[8] of q = 3 belong to the boundary faces).
template <class H1 = NoHook, class H2 = NoHook>
__device__ __forceinline__ void chain_eval(TcShared* S, const float* sbias, const float (&y)[8], float (&nn)[9], int& pd, int wq, int q, long long* twd = nullptr,
                                           H1 hook1 = H1(), H2 hook2 = H2()) {
    const uint32_t tl = S->tmem_base + ((uint32_t)(32 * wq) << 16);     // this warp's TMEM lane window
    // ---- layer-1 A operand: levels 8q..8q+7 -> 4 cells per split at columns 4q
    {
        uint32_t p1[4], p2[4], p3[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) split3(y[2 * i], y[2 * i + 1], p1[i], p2[i], p3[i]);
        tmem_st4(tl + TM_A + 0 * 64 + 4 * q, p1);
        tmem_st4(tl + TM_A + 1 * 64 + 4 * q, p2);
        tmem_st4(tl + TM_A + 2 * 64 + 4 * q, p3);
        wait_st();
        fence_before_sync();
        warp_arrive(&S->bar_a);
    }
    // ---- hidden layers: D -> bias, relu, split -> A operand of the next layer.  Thread (column, quarter q) converts the units
    //      32 s + 8 q .. + 7 in its sub-chunk s = 0..3, so that after sub-chunk s of all four quarters the k-steps 2s, 2s+1 of the next
    //      layer are complete and their MMAs start under the conversion of the following sub-chunk
#pragma unroll 1
    for (int l = 0; l < 2; ++l) {
        if (twd) { const long long a = clock64(); mbar_wait(&S->bar_d, pd); twd[l] += clock64() - a; } else mbar_wait(&S->bar_d, pd);
        pd ^= 1;
        fence_after_sync();
        const float* b = sbias + 128 * l + 8 * q;
        // the whole D slice of this thread first: the next layer's MMAs overwrite D as soon as the first sub-chunks are published
        uint32_t r[32];
#pragma unroll
        for (int sc = 0; sc < 4; ++sc) tmem_ld8(tl + TM_D + 32 * sc + 8 * q, *reinterpret_cast<uint32_t(*)[8]>(&r[8 * sc]));
        wait_ld();
#pragma unroll
        for (int sc = 0; sc < 4; ++sc) {
            uint32_t p1[4], p2[4], p3[4];
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const float4 bb = *reinterpret_cast<const float4*>(b + 32 * sc + 4 * i);
                const float v0 = fmaxf(__uint_as_float(r[8 * sc + 4 * i]) + bb.x, 0.f), v1 = fmaxf(__uint_as_float(r[8 * sc + 4 * i + 1]) + bb.y, 0.f);
                const float v2 = fmaxf(__uint_as_float(r[8 * sc + 4 * i + 2]) + bb.z, 0.f), v3 = fmaxf(__uint_as_float(r[8 * sc + 4 * i + 3]) + bb.w, 0.f);
                split3(v0, v1, p1[2 * i], p2[2 * i], p3[2 * i]);
                split3(v2, v3, p1[2 * i + 1], p2[2 * i + 1], p3[2 * i + 1]);
            }
            tmem_st4(tl + TM_A + 0 * 64 + 16 * sc + 4 * q, p1);
            tmem_st4(tl + TM_A + 1 * 64 + 16 * sc + 4 * q, p2);
            tmem_st4(tl + TM_A + 2 * 64 + 16 * sc + 4 * q, p3);
            wait_st();
            fence_before_sync();
            warp_arrive(sc == 0 ? &S->bar_a : &S->bar_q[sc - 1]);
        }
        if (l == 0) hook1(); else hook2();
    }
    // ---- output layer: NN[8q-1 .. 8q+7]
    if (twd) { const long long a = clock64(); mbar_wait(&S->bar_d, pd); twd[2] += clock64() - a; } else mbar_wait(&S->bar_d, pd);
    pd ^= 1;
    fence_after_sync();
    {
        uint32_t ra[8], rb[8], rc[8], a0 = 0, b0 = 0, c0 = 0;
        tmem_ld8(tl + TM_D + 8 * q, ra);
        tmem_ld8(tl + TM_D + 32 + 8 * q, rb);
        tmem_ld8(tl + TM_D + 64 + 8 * q, rc);
        if (q) { tmem_ld1(tl + TM_D + 8 * q - 1, a0); tmem_ld1(tl + TM_D + 32 + 8 * q - 1, b0); tmem_ld1(tl + TM_D + 64 + 8 * q - 1, c0); }
        wait_ld();
        const float* b3 = sbias + 256;
        // smallest group first: (A B3 + A B2) + A B1
        nn[0] = q ? ((__uint_as_float(c0) + __uint_as_float(b0)) + __uint_as_float(a0)) + b3[8 * q - 1] : 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) nn[i + 1] = ((__uint_as_float(rc[i]) + __uint_as_float(rb[i])) + __uint_as_float(ra[i])) + b3[8 * q + i];
    }
    fence_before_sync();     // the next evaluation's tcgen05.st / MMA may overwrite D and A
}


// =====================================================================================================================
// Split policies.  SplitBF16x3 is the scheme above.  SplitFP16x2 halves the tensor work: x = hi + lo in FP16 (11 + 11
// significand bits), product = lo*hi + hi*lo + hi*hi (dropped lo*lo <= 2^-22), three kind::f16 MMAs per k-step.  FP16 has only 5
// exponent bits, so every operand tensor carries an exact power-of-two scale chosen on the host (nfc_set_weights): weights by
// their max-norm, activations by a bound propagated through the layers from |T_scaled| <= 128 (row-sum norms).  relu commutes
// with positive scaling, so the re-scaling is folded into the bias-add FMA: v' = max(fma(D, c_l, bias'_l), 0) is already the
// scaled activation.  Inputs beyond the bound overflow to inf and are reported as status NaN (never silently wrong).
// =====================================================================================================================
struct SplitBF16x3 {
    static constexpr int NH = 2, CW = 0;                       // default Chain only
    static constexpr int IMG = IMG_BYTES, BIAS_OFF = OFF_B;
    static constexpr float INPUT_BOUND = 3.0e38f;              // BF16 has fp32's exponent range: no bound
    static constexpr bool SPARE_TMEM = false;                  // all 512 tensor-memory columns are in use (three operand splits)
    static __device__ __forceinline__ uint32_t spare(uint32_t tl, int, int q) { return tl + TM_K + 160 + 8 * q; }
    template <int LAYER, int SUB>
    static __device__ __forceinline__ void issue(uint32_t tb, uint32_t img, int nprod) { issue_layer<LAYER, SUB>(tb, img, nprod); }
    template <bool FOLD = true, class H1 = NoHook, class H2 = NoHook>
    static __device__ __forceinline__ void eval(TcShared* S, const float* sb, const float (&y)[8], float (&nn)[9], int& pd, int wq, int q, long long* twd, bool,
                                                H1 h1 = H1(), H2 h2 = H2()) {
        chain_eval(S, sb, y, nn, pd, wq, q, twd, h1, h2);
    }
};

constexpr int W1H_BYTES = 128 * 32 * 2, W2H_BYTES = 128 * 128 * 2, W3H_BYTES = 64 * 128 * 2;   // output layer: [hi; lo] stacked along N (64 rows)
// FP16x2 weight image of a Chain [Conv((CW,1),1=>1,relu)] -> Dense(IN0,128,relu) -> (NH-1) x Dense(128,128,relu) -> Dense(128,31):
// the default Chain (NH = 2, CW = 0), conv-2 / conv-4 (NH = 2, CW = 2 / 4; the first Dense layer's K is zero-padded to 32) and the deeper
// Chain (NH = 3).  (The wider Chain, H = 256, needs 256 KB for its hidden matrix alone and stays on the FP32 SIMT kernels.)
template <int NH_, int CW_>
struct TcLayout {
    static constexpr int NH = NH_, CW = CW_;
    using N = Net<128, NH_, CW_>;
    static constexpr int OFF_W1 = 0;                                          // [hi | lo], K = 32
    static constexpr int OFF_WH = OFF_W1 + 2 * W1H_BYTES;                     // NH - 1 hidden-to-hidden matrices, [hi | lo] each
    static constexpr int OFF_W3 = OFF_WH + (NH_ - 1) * 2 * W2H_BYTES;
    static constexpr int OFF_B = OFF_W3 + W3H_BYTES;                          // fp32: NH x bias'[128], b3[32], consts[16]
    static constexpr int CB = NH_ * 128 + 32;                                 // consts: c_1 .. c_NH, inv3, s_a0, conv w[0..3], conv bias
    static constexpr int C_INV3 = CB + NH_, C_SA0 = CB + NH_ + 1, C_CONVW = CB + NH_ + 2, C_CONVB = CB + NH_ + 6;
    static constexpr int IMG = OFF_B + (CB + 16) * 4;
};
using TcDefault = TcLayout<2, 0>;
constexpr int OFF16_W1 = TcDefault::OFF_W1, OFF16_W2 = TcDefault::OFF_WH, OFF16_W3 = TcDefault::OFF_W3, OFF16_B = TcDefault::OFF_B;
constexpr int IMG16_BYTES = TcDefault::IMG;                                   // 99 520 B
constexpr int IMG16_MAX = TcLayout<3, 0>::IMG;                                // the largest image (deeper Chain): 165 568 B
// exact power-of-two operand scales: s_w[l] of the l-th weight matrix (l = 0 .. NH), s_a[l] of its input (s_a[0]: the network input
// after the Conv front-end, s_a[l]: hidden activations l)
struct Fp16Scales { float s_w[4], s_a[4]; };

__host__ __device__ constexpr uint32_t make_idesc16(int M, int N) {   // kind::f16: A = B = F16 (0), C = F32 (1), both K-major
    return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

template <int NH, int CW>
static __global__ void pack_weights_tc16_kernel(const float* __restrict__ theta, unsigned char* __restrict__ img, const Fp16Scales sc) {
    using L = TcLayout<NH, CW>;
    using N = typename L::N;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int n1 = 128 * 32, n2 = (NH - 1) * 128 * 128, n3 = 32 * 128;
    if (i < n1 + n2 + n3) {
        float w, s;
        int off_hi, off_lo;
        if (i < n1) {
            const int n = i % 128, k = i / 128;
            w = k < N::IN0 ? theta[N::T_W0 + k * 128 + n] : 0.f; s = sc.s_w[0];
            off_hi = L::OFF_W1 + umma_off(n, k, 128); off_lo = off_hi + W1H_BYTES;
        } else if (i < n1 + n2) {
            const int q = i - n1, l = q / (128 * 128), r = q % (128 * 128), n = r % 128, k = r / 128;
            w = theta[N::T_HID + l * N::HID_STRIDE + k * 128 + n]; s = sc.s_w[1 + l];
            off_hi = L::OFF_WH + l * 2 * W2H_BYTES + umma_off(n, k, 128); off_lo = off_hi + W2H_BYTES;
        } else {
            const int q = i - n1 - n2, n = q % 32, k = q / 32;
            w = n < 31 ? theta[N::T_WL + k * 31 + n] : 0.f; s = sc.s_w[NH];
            off_hi = L::OFF_W3 + umma_off(n, k, 64); off_lo = L::OFF_W3 + umma_off(32 + n, k, 64);
        }
        const float ws = w * s;
        const __half hi = __float2half_rn(ws);
        const __half lo = __float2half_rn(ws - __half2float(hi));
        *reinterpret_cast<__half*>(img + off_hi) = hi;
        *reinterpret_cast<__half*>(img + off_lo) = lo;
    } else if (i < n1 + n2 + n3 + L::CB + 16) {
        const int q = i - n1 - n2 - n3;
        float v = 0.f;
        if (q < NH * 128) {                                                                     // bias'_l = b_l * s_a[l]   (l = 1 .. NH)
            const int l = q / 128, j = q % 128;
            v = (l == 0 ? theta[N::T_B0 + j] : theta[N::T_HID + (l - 1) * N::HID_STRIDE + 128 * 128 + j]) * sc.s_a[l + 1];
        } else if (q < L::CB) v = (q - NH * 128 < 31) ? theta[N::T_BL + (q - NH * 128)] : 0.f;   // b3 (the output is un-scaled)
        else if (q < L::CB + NH) { const int l = q - L::CB; v = sc.s_a[l + 1] / (sc.s_a[l] * sc.s_w[l]); }   // c_{l+1}: D -> scaled activation
        else if (q == L::C_INV3) v = 1.f / (sc.s_a[NH] * sc.s_w[NH]);                          // D_out -> NN output
        else if (q == L::C_SA0) v = sc.s_a[0];
        else if (q >= L::C_CONVW && q < L::C_CONVW + 4) v = (q - L::C_CONVW) < CW ? theta[N::T_CONV + (q - L::C_CONVW)] : 0.f;
        else if (q == L::C_CONVB) v = CW ? theta[N::T_CONV + CW] : 0.f;
        reinterpret_cast<float*>(img + L::OFF_B)[q] = v;
    }
}

// (v0, v1) -> packed fp16 (hi) and packed fp16 residual (lo); element 0 in the low half
__device__ __forceinline__ void split2h(float v0, float v1, uint32_t& ph, uint32_t& pl) {
    float f0, f1;
    asm("{\n\t.reg .b16 l, h;\n\tcvt.rn.f16x2.f32 %0, %4, %3;\n\tmov.b32 {l, h}, %0;\n\tcvt.f32.f16 %1, l;\n\tcvt.f32.f16 %2, h;\n\t}"
        : "=r"(ph), "=f"(f0), "=f"(f1) : "f"(v0), "f"(v1));
    float r0, r1;
    sub2(r0, r1, v0, v1, f0, f1);
    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(pl) : "f"(r1), "f"(r0));
}

// the same for a hidden-layer activation relu(v): hi = max(0, rz_f16(v)) and lo = max(0, rn_f16(v - hi)).  Rounding hi towards zero
// makes the residual of a positive v non-negative, so BOTH conversions can carry .relu and the fp32 max(v, 0) disappears from the
// epilogue (4 instead of 5 instructions per activation); for v < 0 both parts are +0.  |lo| < 1 ulp_f16 instead of 1/2: the dropped
// lo*lo term is <= 2^-20 relative (RHS error against float64 stays at the 1e-7 level, bar 1e-5).
__device__ __forceinline__ void split2h_relu(float v0, float v1, uint32_t& ph, uint32_t& pl) {
    float f0, f1;
    asm("{\n\t.reg .b16 l, h;\n\tcvt.rz.relu.f16x2.f32 %0, %4, %3;\n\tmov.b32 {l, h}, %0;\n\tcvt.f32.f16 %1, l;\n\tcvt.f32.f16 %2, h;\n\t}"
        : "=r"(ph), "=f"(f0), "=f"(f1) : "f"(v0), "f"(v1));
    float r0, r1;
    sub2(r0, r1, v0, v1, f0, f1);
    asm("cvt.rn.relu.f16x2.f32 %0, %1, %2;" : "=r"(pl) : "f"(r1), "f"(r0));
}

// LAYER 0: Dense(32, 128) in one go; 1 .. NH-1: hidden-to-hidden; NH: the output layer
template <class L, int LAYER, int SUB>
__device__ __forceinline__ void issue_layer16(uint32_t tmem_base, uint32_t smem_img) {
    constexpr bool OUT = LAYER == L::NH;
    constexpr int Nn = OUT ? 64 : 128;
    constexpr uint32_t woff = LAYER == 0 ? L::OFF_W1 : (OUT ? L::OFF_W3 : L::OFF_WH + (LAYER - 1) * 2 * W2H_BYTES);
    constexpr uint32_t wbytes = LAYER == 0 ? W1H_BYTES : W2H_BYTES;
    constexpr uint32_t idesc = make_idesc16(128, Nn);
    constexpr uint32_t kstep = (2u * Nn * 16u) >> 4;
    const uint32_t d = tmem_base + TM_D;
    const uint64_t desc0 = make_bdesc(smem_img + woff, Nn);
    const uint32_t lo0 = (uint32_t)desc0, hi0 = (uint32_t)(desc0 >> 32);
    const uint32_t a_base = tmem_base + TM_A;
    if (OUT) {                   // stacked [B_hi; B_lo]: D[:, 0:32] + D[:, 32:64] = A (B_hi + B_lo), a split lo first
#pragma unroll
        for (int sA = 1; sA >= 0; --sA)
#pragma unroll
            for (int j = 2 * SUB; j < 2 * SUB + 2; ++j) {
                if (sA == 1 && j == 0) mma_ts2<false>(d, a_base + sA * 64 + 8 * j, lo0 + kstep * j, hi0, idesc);
                else mma_ts2<true>(d, a_base + sA * 64 + 8 * j, lo0 + kstep * j, hi0, idesc);
            }
    } else {
        constexpr int sa[3] = {1, 0, 0};      // (a split, b split): lo*hi, hi*lo, hi*hi
        constexpr int sb[3] = {0, 1, 0};
#pragma unroll
        for (int p = 0; p < 3; ++p)
#pragma unroll
            for (int j = 2 * SUB; j < 2 * SUB + 2; ++j) {
                const uint32_t blo = lo0 + ((sb[p] * wbytes) >> 4) + kstep * j;
                if (p == 0 && j == 0) mma_ts2<false>(d, a_base + sa[p] * 64 + 8 * j, blo, hi0, idesc);
                else mma_ts2<true>(d, a_base + sa[p] * 64 + 8 * j, blo, hi0, idesc);
            }
    }
}

// FP16x2 version of chain_eval; sb points at bias1' (then bias2', bias3, consts).  FOLD: relu folded into the conversions (split2h_relu);
// the RECORD pass of loss_grad keeps the round-to-nearest split so that its trajectories stay those the gradient tests were pinned with
// y: the network input of this thread's 8 levels (after the Conv front-end, if the Chain has one)
template <class L, bool FOLD, class H1 = NoHook, class H2 = NoHook>
__device__ __forceinline__ void chain_eval16(TcShared* S, const float* sb, const float (&y)[8], float (&nn)[9], int& pd, int wq, int q, long long* twd,
                                             H1 hook1 = H1(), H2 hook2 = H2()) {
    const uint32_t tl = S->tmem_base + ((uint32_t)(32 * wq) << 16);
    const float s_a0 = sb[L::C_SA0];
    {
        uint32_t ph[4], pl[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float s0, s1;
            mul2s(s0, s1, y[2 * i], y[2 * i + 1], s_a0);
            split2h(s0, s1, ph[i], pl[i]);
        }
        tmem_st4(tl + TM_A + 0 * 64 + 4 * q, ph);
        tmem_st4(tl + TM_A + 1 * 64 + 4 * q, pl);
        wait_st();
        fence_before_sync();
        warp_arrive(&S->bar_a);
    }
    // hidden layers: thread (column, quarter q) converts the units 32 s + 8 q .. + 7 in its sub-chunk s = 0..3 (see chain_eval)
#pragma unroll 1
    for (int l = 0; l < L::NH; ++l) {
        if (twd) { const long long a = clock64(); mbar_wait(&S->bar_d, pd); twd[l] += clock64() - a; } else mbar_wait(&S->bar_d, pd);
        pd ^= 1;
        fence_after_sync();
        const float* b = sb + 128 * l + 8 * q;
        const float cl = sb[L::CB + l];
        uint32_t r[32];
#pragma unroll
        for (int sc = 0; sc < 4; ++sc) tmem_ld8(tl + TM_D + 32 * sc + 8 * q, *reinterpret_cast<uint32_t(*)[8]>(&r[8 * sc]));
        wait_ld();
#pragma unroll
        for (int sc = 0; sc < 4; ++sc) {
            uint32_t ph[4], pl[4];
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const float4 bb = *reinterpret_cast<const float4*>(b + 32 * sc + 4 * i);
                float v0, v1, v2, v3;
                fma2s(v0, v1, __uint_as_float(r[8 * sc + 4 * i]), __uint_as_float(r[8 * sc + 4 * i + 1]), cl, bb.x, bb.y);
                fma2s(v2, v3, __uint_as_float(r[8 * sc + 4 * i + 2]), __uint_as_float(r[8 * sc + 4 * i + 3]), cl, bb.z, bb.w);
                if (FOLD) {
                    split2h_relu(v0, v1, ph[2 * i], pl[2 * i]);
                    split2h_relu(v2, v3, ph[2 * i + 1], pl[2 * i + 1]);
                } else {
                    split2h(fmaxf(v0, 0.f), fmaxf(v1, 0.f), ph[2 * i], pl[2 * i]);
                    split2h(fmaxf(v2, 0.f), fmaxf(v3, 0.f), ph[2 * i + 1], pl[2 * i + 1]);
                }
            }
            tmem_st4(tl + TM_A + 0 * 64 + 16 * sc + 4 * q, ph);
            tmem_st4(tl + TM_A + 1 * 64 + 16 * sc + 4 * q, pl);
            wait_st();
            fence_before_sync();
            warp_arrive(sc == 0 ? &S->bar_a : &S->bar_q[sc - 1]);
        }
        if (l == 0) hook1(); else if (l == L::NH - 1) hook2();
    }
    if (twd) { const long long a = clock64(); mbar_wait(&S->bar_d, pd); twd[L::NH] += clock64() - a; } else mbar_wait(&S->bar_d, pd);
    pd ^= 1;
    fence_after_sync();
    {
        uint32_t ra[8], rb[8], a0 = 0, b0 = 0;
        tmem_ld8(tl + TM_D + 8 * q, ra);
        tmem_ld8(tl + TM_D + 32 + 8 * q, rb);
        if (q) { tmem_ld1(tl + TM_D + 8 * q - 1, a0); tmem_ld1(tl + TM_D + 32 + 8 * q - 1, b0); }
        wait_ld();
        const float* b3 = sb + 128 * L::NH;
        const float inv3 = sb[L::C_INV3];
        nn[0] = q ? fmaf(__uint_as_float(b0) + __uint_as_float(a0), inv3, b3[8 * q - 1]) : 0.f;
#pragma unroll
        for (int i = 0; i < 8; i += 2) {
            float t0, t1;
            add2(t0, t1, __uint_as_float(rb[i]), __uint_as_float(rb[i + 1]), __uint_as_float(ra[i]), __uint_as_float(ra[i + 1]));
            fma2s(nn[i + 1], nn[i + 2], t0, t1, inv3, b3[8 * q + i], b3[8 * q + i + 1]);
        }
    }
    fence_before_sync();
}

template <int NH_, int CW_>
struct SplitFP16x2T {
    using L = TcLayout<NH_, CW_>;
    static constexpr int NH = NH_, CW = CW_;
    static constexpr int IMG = L::IMG, BIAS_OFF = L::OFF_B;
    template <int LAYER, int SUB>
    static __device__ __forceinline__ void issue(uint32_t tb, uint32_t img, int) { issue_layer16<L, LAYER, SUB>(tb, img); }
    static constexpr float INPUT_BOUND = 128.f;
    // free 32-column vector slots in tensor memory: the third operand-split region (two slots) and the k7 slot (k7 stays in registers)
    static constexpr bool SPARE_TMEM = true;
    static __device__ __forceinline__ uint32_t spare(uint32_t tl, int i, int q) { return tl + (i < 2 ? TM_A + 128 + 32 * i : TM_K + 160) + 8 * q; }
    // The operand scales assume |T_scaled| <= 128 (the hidden-layer bounds then hold rigorously).  The network input is SATURATED
    // at the bound: inside the adaptive solver only wildly wrong trial steps get there and the error estimate rejects them exactly
    // as in FP32 (the flux divergence itself uses the unclamped state); the solver flags a column whose ACCEPTED state leaves the
    // range (status NaN).  With `poison` (nfc_rhs: no controller) an out-of-range or NaN input yields NaN output instead.
    template <bool FOLD = true, class H1 = NoHook, class H2 = NoHook>
    static __device__ __forceinline__ void eval(TcShared* S, const float* sb, const float (&y)[8], float (&nn)[9], int& pd, int wq, int q, long long* twd,
                                                bool poison, H1 h1 = H1(), H2 h2 = H2()) {
        float yc[8];
        bool bad = false;
        // with a Conv front-end `y` is its output, whose bound (128 sum|w| + |b|) went into the scale s_a0: saturate where the scaled operand
        // reaches 16384
        const float bound = CW ? 16384.f / sb[L::C_SA0] : INPUT_BOUND;
#pragma unroll
        for (int i = 0; i < 8; ++i) { yc[i] = fminf(fmaxf(y[i], -bound), bound); bad |= !(fabsf(y[i]) <= bound); }
        chain_eval16<L, FOLD>(S, sb, yc, nn, pd, wq, q, twd, h1, h2);
        if (poison && bad) {
#pragma unroll
            for (int i = 0; i < 9; ++i) nn[i] = __int_as_float(0x7fc00000);
        }
    }
};
using SplitFP16x2 = SplitFP16x2T<2, 0>;      // the default Chain

// f (8 levels of this thread) from the Chain output and the stage input: faces k = 8q .. 8q+8,
// F_0 = bottom, F_32 = top, else NN[k-1] - min(0, K_CA (y_k - y_{k-1}) / dz);  f_k = -pref (F_{k+1} - F_k) / dz
// ylo = y_{8q-1} (from quarter q-1), yhi = y_{8q+8} (from quarter q+1)
__device__ __forceinline__ void faces_to_f(const float (&y)[8], float ylo, float yhi, const float (&nn)[9], float bot, float top, float kk,
                                           float pref, int q, float (&f)[8]) {
    float Fprev = 0.f;
#pragma unroll
    for (int i = 0; i <= 8; ++i) {
        const int k = 8 * q + i;
        const float yk = i < 8 ? y[i < 8 ? i : 0] : yhi;
        const float ykm = i > 0 ? y[i > 0 ? i - 1 : 0] : ylo;
        float F = nn[i] - fminf(kk * (yk - ykm), 0.f);
        if (k == 0) F = bot;
        if (k == NZ) F = top;
        if (i > 0) f[i - 1] = -pref * ((F - Fprev) * (float)NZ);
        Fprev = F;
    }
}

__device__ __forceinline__ void cta_bar() { asm volatile("bar.sync 1, 512;" ::: "memory"); }      // the 512 column threads

// Conv((CW,1),1=>1,relu) front-end of the conv-2 / conv-4 Chains on this thread's 8 levels (train_free_convection_nde.jl:138-153; a true
// convolution: x_i = relu(b + sum_j w_j y_{i+CW-1-j})).  The CW-1 levels beyond this quarter come from the next quarter's thread through
// shared memory (one barrier of the column threads per evaluation); outputs beyond Nz-CW+1 are zero (their rows of the first Dense layer are
// zero padding).  cw: w[0..3], bias at cw[4] (consts of the weight image).
template <int CW>
__device__ __forceinline__ void conv_front(TcShared* S, const float* cw, const float (&y)[8], float (&x)[8], int xb, int c, int q) {
#pragma unroll
    for (int j = 0; j < CW - 1; ++j) S->hal[xb][c][q][j] = y[j];
    cta_bar();
    float ext[8 + 3];
#pragma unroll
    for (int i = 0; i < 8; ++i) ext[i] = y[i];
#pragma unroll
    for (int j = 0; j < 3; ++j) ext[8 + j] = (j < CW - 1 && q < 3) ? S->hal[xb][c][q + 1][j < CW - 1 ? j : 0] : 0.f;
    const float b = cw[4];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        float acc = 0.f;
#pragma unroll
        for (int j = 0; j < CW; ++j) acc = fmaf(cw[j], ext[i + CW - 1 - j], acc);
        x[i] = (8 * q + i < NZ - CW + 1) ? fmaxf(acc + b, 0.f) : 0.f;
    }
}

// MMA warp main loop: three layers per Chain evaluation until the column threads raise `stop`.  Layer 1 (K = 32) is issued in one go;
// the hidden-to-hidden and the output layer in four sub-batches (k-steps 2s, 2s+1) as the column threads publish their sub-chunks.
template <class SP, int LAYER>
__device__ __forceinline__ void mma_issue_subs(TcShared* S, uint32_t tb0, uint32_t smem_img, int& pa, int (&pq)[3], long long& t_wait) {
    const long long c0 = clock64();
    mbar_wait(&S->bar_a, pa); pa ^= 1;
    fence_after_sync();
    t_wait += clock64() - c0;
    // The operand addresses must be re-derived inside the loop (an add on a uniform register per MMA): as loop invariants the
    // compiler hoists them all and, with the 32 registers this warp keeps, spills them -- a local-memory load in front of the MMA.
    uint32_t tb = tb0, img = smem_img;
    asm volatile("" : "+r"(tb), "+r"(img));
    SP::template issue<LAYER, 0>(tb, img, S->nprod);
    mbar_wait(&S->bar_q[0], pq[0]); pq[0] ^= 1;
    fence_after_sync();
    asm volatile("" : "+r"(tb), "+r"(img));
    SP::template issue<LAYER, 1>(tb, img, S->nprod);
    mbar_wait(&S->bar_q[1], pq[1]); pq[1] ^= 1;
    fence_after_sync();
    asm volatile("" : "+r"(tb), "+r"(img));
    SP::template issue<LAYER, 2>(tb, img, S->nprod);
    mbar_wait(&S->bar_q[2], pq[2]); pq[2] ^= 1;
    fence_after_sync();
    asm volatile("" : "+r"(tb), "+r"(img));
    SP::template issue<LAYER, 3>(tb, img, S->nprod);
    mma_commit(&S->bar_d);
}

template <class SP>
__device__ __forceinline__ void mma_warp_loop(TcShared* S, uint32_t smem_img, int lane, long long* dbg = nullptr) {
    if (lane == 0) {
        int pa = 0, pq[3] = {0, 0, 0};
        long long tw[3] = {0, 0, 0}, ti[3] = {0, 0, 0}, n = 0;
        const uint32_t tb0 = S->tmem_base;          // read once; the empty asm statements below keep the derived operand addresses inside the loop
        while (true) {
            long long c0 = clock64();
            mbar_wait(&S->bar_a, pa); pa ^= 1;
            if (*reinterpret_cast<volatile int*>(&S->stop)) break;
            fence_after_sync();
            long long c1 = clock64();
            uint32_t tb = tb0, img = smem_img;
            asm volatile("" : "+r"(tb), "+r"(img));
            SP::template issue<0, 0>(tb, img, S->nprod);
            mma_commit(&S->bar_d);
            long long c2 = clock64();
            tw[0] += c1 - c0; ti[0] += c2 - c1;
            long long w = 0;
            mma_issue_subs<SP, 1>(S, tb0, smem_img, pa, pq, w);
            if constexpr (SP::NH == 3) mma_issue_subs<SP, 2>(S, tb0, smem_img, pa, pq, w);      // third hidden layer (deeper Chain)
            c1 = clock64(); tw[1] += w; ti[1] += c1 - c2 - w;
            w = 0;
            mma_issue_subs<SP, SP::NH>(S, tb0, smem_img, pa, pq, w);                            // output layer
            c2 = clock64(); tw[2] += w; ti[2] += c2 - c1 - w;
            n += 1;
        }
        if (dbg && blockIdx.x == 0)      // per layer: waiting for the first sub-chunk | everything up to the commit (issue + later sub-chunks)
            for (int l = 0; l < 3; ++l) { dbg[5 * l] = tw[l]; dbg[5 * l + 1] = ti[l]; dbg[5 * l + 2] = 0; dbg[5 * l + 3] = 0; dbg[5 * l + 4] = n; }
    }
    __syncwarp();
}

// common prologue: barriers, TMEM allocation, weight image.  Returns after everything is visible to all threads.
__device__ __forceinline__ void tc_prologue(TcShared* S, unsigned char* simg, const unsigned char* gimg, int warp, int img_bytes, int nprod = 6) {
    if (threadIdx.x == 0) {
        mbar_init(&S->bar_w, 1);
        mbar_init(&S->bar_a, NCW);
        mbar_init(&S->bar_q[0], NCW); mbar_init(&S->bar_q[1], NCW); mbar_init(&S->bar_q[2], NCW);
        mbar_init(&S->bar_d, 1);
        S->stop = 0;
        S->nprod = nprod;
    }
    __syncthreads();
    if (warp == NCT / 32) tmem_alloc(&S->tmem_base, TM_COLS);
    if (threadIdx.x == 0) {
        mbar_expect_tx(&S->bar_w, img_bytes);
        for (unsigned off = 0; off < (unsigned)img_bytes; off += 32768u) {
            const unsigned n = img_bytes - off < 32768u ? img_bytes - off : 32768u;
            bulk_g2s(simg + off, gimg + off, n, &S->bar_w);
        }
    }
    mbar_wait(&S->bar_w, 0);
    // generic-proxy view of the image is not needed by the MMA (async proxy reads smem); biases are read by LDS after this wait
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
}

// ---- test / nfc_rhs kernel: f(T) for whole tiles of 128 columns ---------------------------------------------------------
template <class SP>
__global__ void __launch_bounds__(NTHREADS, 1)
rhs_tc_kernel(const unsigned char* __restrict__ gimg, const float* __restrict__ T, const float* __restrict__ colp, float* __restrict__ out,
              long long B, float pref, int use_kca) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ TcShared S;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    tc_prologue(&S, smem_raw, gimg, warp, SP::IMG);
    const uint32_t smem_img = smem_u32(smem_raw);
    const float* sbias = reinterpret_cast<const float*>(smem_raw + SP::BIAS_OFF);
    const long long ntiles = (B + TILE - 1) / TILE;
    if (warp >= NCT / 32) {
        reg_dec<REGS_MMA>();
        if (warp == NCT / 32) mma_warp_loop<SP>(&S, smem_img, lane);
    } else {
        reg_inc<REGS_COL>();
        const int wq = warp & 3, q = warp >> 2, c = 32 * wq + lane;
        int pd = 0, xb = 0;
        for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            const long long col = tile * TILE + c;
            const bool ok = col < B;
            float y[8], nn[9], f[8];
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const float4 v = ok ? *reinterpret_cast<const float4*>(T + col * NZ + 8 * q + 4 * i) : make_float4(0.f, 0.f, 0.f, 0.f);
                y[4 * i] = v.x; y[4 * i + 1] = v.y; y[4 * i + 2] = v.z; y[4 * i + 3] = v.w;
            }
            xb ^= 1;
            S.xch[xb][c][q][0] = y[0];
            S.xch[xb][c][q][1] = y[7];
            if constexpr (SP::CW > 0) {
                float xin[8];
                bool bad = false;
#pragma unroll
                for (int i = 0; i < 8; ++i) bad |= !(fabsf(y[i]) <= SP::INPUT_BOUND);
                conv_front<SP::CW>(&S, sbias + SP::L::C_CONVW, y, xin, xb, c, q);
                SP::eval(&S, sbias, xin, nn, pd, wq, q, nullptr, false);
                if (bad) {
#pragma unroll
                    for (int i = 0; i < 9; ++i) nn[i] = __int_as_float(0x7fc00000);
                }
            } else {
                SP::eval(&S, sbias, y, nn, pd, wq, q, nullptr, true);   // its mbarrier traffic orders the exchange (double buffered)
            }
            const float bot = ok ? colp[col * 3 + 0] : 0.f, top = ok ? colp[col * 3 + 1] : 0.f;
            const float kk = (ok && use_kca) ? colp[col * 3 + 2] * (float)NZ : 0.f;
            const float ylo = q > 0 ? S.xch[xb][c][q - 1][1] : 0.f, yhi = q < 3 ? S.xch[xb][c][q + 1][0] : 0.f;
            faces_to_f(y, ylo, yhi, nn, bot, top, kk, pref, q, f);
            if (ok) {
                *reinterpret_cast<float4*>(out + col * NZ + 8 * q) = make_float4(f[0], f[1], f[2], f[3]);
                *reinterpret_cast<float4*>(out + col * NZ + 8 * q + 4) = make_float4(f[4], f[5], f[6], f[7]);
            }
        }
        // release the MMA warp
        asm volatile("bar.sync 1, 512;" ::: "memory");
        if (threadIdx.x == 0) S.stop = 1;
        asm volatile("bar.sync 1, 512;" ::: "memory");
        warp_arrive(&S.bar_a);
    }
    fence_before_sync();
    __syncthreads();
    if (warp == NCT / 32) tmem_dealloc(S.tmem_base, TM_COLS);
}

}  // namespace tc
}  // namespace nfc

// =====================================================================================================================
// Persistent adaptive Tsit5 solve on the tensor-core Chain: one CTA per SM, 128 column slots (two threads per column),
// every slot with its own PI controller, finished slots refilled from the global queue (same semantics as
// nfc_forward.cuh:solve_kernel; src/solve.jl:4).
// =====================================================================================================================
#include "nfc_forward.cuh"

namespace nfc {
namespace tc {

// k_j (j = 2..7) chunk of this thread in tensor memory: column TM_K + 32 (j-2) + 8 q, 8 cells
__device__ __forceinline__ uint32_t kaddr(uint32_t tl, int j, int q) { return tl + TM_K + 32 * (j - 2) + 8 * q; }

// step-boundary arithmetic (fast_div, fast_sqrt, err_term, pi_controller): nfc_device.cuh, shared by every adaptive kernel
using nfc::err_term;
using nfc::fast_div;
using nfc::fast_sqrt;
using nfc::pi_controller;
constexpr int SAVE_SMEM = 2048;           // save times kept in shared memory by the one-tile kernel (longer lists are read from global memory)

// acc[i] = sum_j coef[j-1] * k_j[i] over j = 1..nk (k1 in registers, k2..k6 from tensor memory), one fma chain in the order of j.
// nk is uniform over the CTA (tcgen05.ld is warp-collective); all loads are in flight before the first use
__device__ __forceinline__ void combine_n(uint32_t tl, int q, const float (&k1)[8], const float* coef, int nk, float (&acc)[8]) {
    float ka[8], kb[8], kc[8], kd[8], ke[8];
    if (nk >= 2) tmem_ld8f(kaddr(tl, 2, q), ka);
    if (nk >= 3) tmem_ld8f(kaddr(tl, 3, q), kb);
    if (nk >= 4) tmem_ld8f(kaddr(tl, 4, q), kc);
    if (nk >= 5) tmem_ld8f(kaddr(tl, 5, q), kd);
    if (nk >= 6) tmem_ld8f(kaddr(tl, 6, q), ke);
    scal8(acc, coef[0], k1);
    if (nk >= 2) { wait_ld(); axpy8(acc, coef[1], ka); }
    if (nk >= 3) axpy8(acc, coef[2], kb);
    if (nk >= 4) axpy8(acc, coef[3], kc);
    if (nk >= 5) axpy8(acc, coef[4], kd);
    if (nk >= 6) axpy8(acc, coef[5], ke);
}

// Tsit5 dense output in power basis, u(t + th h) = u + h th (k1 + th (c2 + th (c3 + th c4))): the k1..k6 part of c2, c3, c4 (each one fma
// chain in the order of j; every stage vector is loaded once for the three); the k7 terms are added by dense_k7
__device__ __forceinline__ void dense_partial(uint32_t tl, int q, const float (&k1)[8], float (&c2)[8], float (&c3)[8], float (&c4)[8]) {
    const float r2[6] = {-2.763706197274826f, 0.13169999999999998f, 3.9302962368947516f, -12.411077166933676f, 37.50931341651104f, -27.896526289197286f};
    const float r3[6] = {2.9132554618219126f, -0.2234f, -5.941033872131505f, 30.33818863028232f, -88.1789048947664f, 65.09189467479366f};
    const float r4[6] = {-1.0530884977290216f, 0.1017f, 2.490627285651253f, -16.548102889244902f, 47.37952196281928f, -34.87065786149661f};
    float ka[8], kb[8];
    tmem_ld8f(kaddr(tl, 2, q), ka);
    tmem_ld8f(kaddr(tl, 3, q), kb);
    scal8(c2, r2[0], k1); scal8(c3, r3[0], k1); scal8(c4, r4[0], k1);
    wait_ld();
    axpy8(c2, r2[1], ka); axpy8(c3, r3[1], ka); axpy8(c4, r4[1], ka);
    axpy8(c2, r2[2], kb); axpy8(c3, r3[2], kb); axpy8(c4, r4[2], kb);
    tmem_ld8f(kaddr(tl, 4, q), ka);
    tmem_ld8f(kaddr(tl, 5, q), kb);
    wait_ld();
    axpy8(c2, r2[3], ka); axpy8(c3, r3[3], ka); axpy8(c4, r4[3], ka);
    axpy8(c2, r2[4], kb); axpy8(c3, r3[4], kb); axpy8(c4, r4[4], kb);
    tmem_ld8f(kaddr(tl, 6, q), ka);
    wait_ld();
    axpy8(c2, r2[5], ka); axpy8(c3, r3[5], ka); axpy8(c4, r4[5], ka);
}
__device__ __forceinline__ void dense_k7(const float (&k7)[8], float (&c2)[8], float (&c3)[8], float (&c4)[8]) {
    axpy8(c2, 1.5f, k7); axpy8(c3, -4.0f, k7); axpy8(c4, 2.5f, k7);
}
// v = u + h th (k1 + th (c2 + th (c3 + th c4)))  (the dense output at th in [0, 1])
__device__ __forceinline__ void dense_eval(float (&v)[8], float hs, float th, const float (&u)[8], const float (&k1)[8], const float (&c2)[8],
                                           const float (&c3)[8], const float (&c4)[8]) {
#pragma unroll
    for (int i = 0; i < 8; i += 2) {
        float a0, a1;
        fma2s(a0, a1, c4[i], c4[i + 1], th, c3[i], c3[i + 1]);
        fma2s(a0, a1, a0, a1, th, c2[i], c2[i + 1]);
        fma2s(a0, a1, a0, a1, th, k1[i], k1[i + 1]);
        mul2s(a0, a1, a0, a1, th);
        fma2s(v[i], v[i + 1], a0, a1, hs, u[i], u[i + 1]);
    }
}

// RECORD: forward pass of nfc_loss_grad (nfc_adjoint.cuh): instead of the trajectory it writes the residuals at the save points, the
// dense tape (u_n and the interpolant in power basis per accepted step; the row index is the column's accepted-step count) and the
// sum of squared residuals -- the same quantities as solve_kernel<RECORD>.
template <class SP, bool RECORD = false>
__global__ void __launch_bounds__(NTHREADS, 1) solve_tc_kernel(const unsigned char* __restrict__ gimg, const SolveParams P) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ TcShared S;
    __shared__ float s_save[SAVE_SMEM];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float* const sv = P.n_save <= SAVE_SMEM ? s_save : P.saveat;       // read by every thread at every accepted step
    if (P.n_save <= SAVE_SMEM) for (int i = threadIdx.x; i < P.n_save; i += NTHREADS) s_save[i] = P.saveat[i];      // visible after the prologue's __syncthreads
#ifdef NFC_DEBUG_TIMERS
    tc_prologue(&S, smem_raw, gimg, warp, SP::IMG, (!RECORD && P.tape_cap >= 1 && P.tape_cap <= 6) ? P.tape_cap : 6);   // debug build: tape_cap doubles as a timing knob
    long long* const dbgbuf = RECORD ? nullptr : reinterpret_cast<long long*>(P.resid);                                  //              P.resid carries the phase timers
#else
    tc_prologue(&S, smem_raw, gimg, warp, SP::IMG);
    long long* const dbgbuf = nullptr;
#endif
    const uint32_t smem_img = smem_u32(smem_raw);
    const float* sbias = reinterpret_cast<const float*>(smem_raw + SP::BIAS_OFF);
    if (warp >= NCT / 32) {
        reg_dec<REGS_MMA>();
        if (warp == NCT / 32) mma_warp_loop<SP>(&S, smem_img, lane, dbgbuf);
    } else {
        reg_inc<REGS_COL>();
        const int wq = warp & 3, q = warp >> 2, c = 32 * wq + lane;
        long long twd[4] = {0, 0, 0, 0}, tstage = 0, nsteps = 0, tb[6] = {0, 0, 0, 0, 0, 0}, tmark = 0;     // tstage: cycles inside the six-stage loop; the rest of the kernel's time is step boundaries
        long long* const twdp = dbgbuf ? twd : nullptr;      // phase timers: debug build with NFC_TC_DEBUG=1 only
        const long long tstart = clock64();
        const uint32_t tl = S.tmem_base + ((uint32_t)(32 * wq) << 16);
        int pd = 0, xb = 0, pb = 0;
        float u[8], k1[8], y[8];
        int col = -1, si = 0, nacc = 0, nrej = 0, status = ST_OK;
        float t = 0.f, dt = 0.f, qold = 1e-4f, bot = 0.f, top = 0.f, kk = 0.f;
        bool queue_empty = false, done = true;       // done => slot wants a (new) column
        double lsum = 0.0;                           // RECORD: this thread's sum of squared residuals
        const float zero8[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int i = 0; i < 8; ++i) { u[i] = 0.f; k1[i] = 0.f; }
        while (true) {
            if (twdp) { const long long n = clock64(); if (tmark) tb[3] += n - tmark; tmark = n; }
            // ---------------- retire / refill ----------------
            if (q == 0) {
                int nc = -2;                                    // -2: keep the current column
                if (done) {
                    if (col >= 0) {
                        if (P.naccept) P.naccept[col] = nacc;
                        if (P.nreject) P.nreject[col] = nrej;
                        if (P.status) P.status[col] = status;
                        if (P.work) P.work[col] = nacc + nrej;
                        if (RECORD) P.tape_len[col] = min(nacc, tape_capacity(P, col));
                        atomicAdd(&P.counters[CTR_ACCEPT], (unsigned long long)nacc);
                        atomicAdd(&P.counters[CTR_REJECT], (unsigned long long)nrej);
                    }
                    nc = -1;
                    if (!queue_empty) {
                        const long long nx = (long long)atomicAdd(&P.counters[CTR_QUEUE], 1ULL);
                        if (nx < P.B) nc = P.order ? P.order[nx] : (int)nx; else queue_empty = true;
                    }
                }
                S.newcol[c] = nc;
            }
            if (threadIdx.x == 0) S.any = 0;
            cta_bar();
            if (twdp) { const long long n = clock64(); tb[0] += n - tmark; tmark = n; }
            {
                const int nc = S.newcol[c];
                // tcgen05.ld/st are warp-collective: NEVER issue them under a lane-dependent branch.  k2..k7 are dead at a
                // step boundary, so when any lane of the warp changes column the whole warp's k region is cleared
                // (stale stage vectors meet zero tableau entries and must stay finite).
                if (__any_sync(FULL, nc != -2)) {
#pragma unroll
                    for (int j = 2; j <= 7; ++j) tmem_st8f(kaddr(tl, j, q), zero8);
                    wait_st();
                }
                if (nc != -2) {
                    col = nc; done = false; nacc = 0; nrej = 0; status = ST_OK; t = 0.f; qold = 1e-4f; si = 0;
                    if (nc >= 0) {
                        const long long o = (long long)nc * NZ + 8 * q;
#pragma unroll
                        for (int i = 0; i < 2; ++i) {
                            const float4 a = *reinterpret_cast<const float4*>(P.T0 + o + 4 * i), b = *reinterpret_cast<const float4*>(P.k1_init + o + 4 * i);
                            u[4 * i] = a.x; u[4 * i + 1] = a.y; u[4 * i + 2] = a.z; u[4 * i + 3] = a.w;
                            k1[4 * i] = b.x; k1[4 * i + 1] = b.y; k1[4 * i + 2] = b.z; k1[4 * i + 3] = b.w;
                        }
                        dt = P.dt_init[nc];
                        bot = P.colp[(long long)nc * 3 + 0]; top = P.colp[(long long)nc * 3 + 1];
                        kk = P.use_kca ? P.colp[(long long)nc * 3 + 2] * (float)NZ : 0.f;
                        while (si < P.n_save && sv[si] <= 0.f) {
                            const long long o2i = ((long long)nc * P.n_save + si) * NZ + 8 * q;
                            if (RECORD) {
                                const float4 ta = *reinterpret_cast<const float4*>(P.Ttrue + o2i), tb = *reinterpret_cast<const float4*>(P.Ttrue + o2i + 4);
                                const float r[8] = {u[0] - ta.x, u[1] - ta.y, u[2] - ta.z, u[3] - ta.w, u[4] - tb.x, u[5] - tb.y, u[6] - tb.z, u[7] - tb.w};
                                *reinterpret_cast<float4*>(P.resid + o2i) = make_float4(r[0], r[1], r[2], r[3]);
                                *reinterpret_cast<float4*>(P.resid + o2i + 4) = make_float4(r[4], r[5], r[6], r[7]);
                                float sse = 0.f;
#pragma unroll
                                for (int i = 0; i < 8; ++i) sse = fmaf(r[i], r[i], sse);
                                lsum += (double)sse;
                            } else if (P.Tout) {                   // null: counting pass of the compact tape (nfc_adjoint.cuh)
                                float* o2 = P.Tout + o2i;          // streaming stores: 1 GB of trajectories must not evict the inputs of later columns from L2
                                __stcs(reinterpret_cast<float4*>(o2), make_float4(u[0], u[1], u[2], u[3]));
                                __stcs(reinterpret_cast<float4*>(o2 + 4), make_float4(u[4], u[5], u[6], u[7]));
                            }
                            ++si;
                        }
                    } else {
                        queue_empty = true;
                        dt = 0.f; bot = 0.f; top = 0.f; kk = 0.f;
#pragma unroll
                        for (int i = 0; i < 8; ++i) { u[i] = 0.f; k1[i] = 0.f; }
                    }
                }
                if (col >= 0) S.any = 1;
            }
            cta_bar();
            if (twdp) { const long long n = clock64(); tb[1] += n - tmark; tmark = n; }
            if (!S.any) break;
            // ---------------- one attempted Tsit5 step ----------------
            float hs = 0.f;
            bool last = false;
            if (col >= 0) {
                hs = fminf(dt, P.tf - t);
                if (t + hs >= P.tf * (1.0f - 1e-6f)) { hs = P.tf - t; last = true; }
            }
            // The linear combination a stage needs, sum_j a_sj k_j, is formed ahead of time: its k1 .. k_{s+1} part while the tensor pipe runs
            // layer 2 of stage s (hook of the Chain evaluation), its last term when f(y_s) = k_{s+2} arrives.  After the last stage the hooks
            // form the error estimate's and the dense output's k1..k6 parts instead.  Same fma chains, term by term, as a combination
            // computed in one go.
            float k7[8], pre[8];
            const long long tloop = twdp ? clock64() : 0;
            scal8(pre, c_tsit5_a[0][0], k1);
            const float bt[6] = {NFC_BT1, NFC_BT2, NFC_BT3, NFC_BT4, NFC_BT5, NFC_BT6};
#pragma unroll 1
            for (int s = 0; s < 6; ++s) {
#pragma unroll
                for (int i = 0; i < 8; i += 2) fma2s(y[i], y[i + 1], pre[i], pre[i + 1], hs, u[i], u[i + 1]);
                xb ^= 1;
                S.xch[xb][c][q][0] = y[0];
                S.xch[xb][c][q][1] = y[7];
                float nn[9], f[8];
                auto h1 = [&]() {       // two calls: the coefficient tables live in different address spaces (constant bank / immediates)
                    if (s < 5) combine_n(tl, q, k1, c_tsit5_a[s + 1], s + 1, pre);
                    else combine_n(tl, q, k1, bt, 6, pre);
                };
                auto h2 = [&]() {
                    if (SP::SPARE_TMEM && s == 5) {      // dense-output partial sums, parked in free tensor-memory columns until the step is judged
                        float d2[8], d3[8], d4[8];
                        dense_partial(tl, q, k1, d2, d3, d4);
                        tmem_st8f(SP::spare(tl, 0, q), d2);
                        tmem_st8f(SP::spare(tl, 1, q), d3);
                        tmem_st8f(SP::spare(tl, 2, q), d4);
                        wait_st();
                    }
                };
                float xin[8];
                if constexpr (SP::CW > 0) conv_front<SP::CW>(&S, sbias + SP::L::C_CONVW, y, xin, xb, c, q);
                const float (&xi)[8] = SP::CW > 0 ? xin : y;
                SP::template eval<true>(&S, sbias, xi, nn, pd, wq, q, twdp, false, h1, h2);      // nfc_solve and the RECORD pass of nfc_loss_grad share ONE arithmetic
                const float ylo = q > 0 ? S.xch[xb][c][q - 1][1] : 0.f, yhi = q < 3 ? S.xch[xb][c][q + 1][0] : 0.f;
                faces_to_f(y, ylo, yhi, nn, bot, top, kk, P.pref, q, f);
                if (s < 5) {
                    tmem_st8f(kaddr(tl, s + 2, q), f);       // completed by the wait::st of the next evaluation's operand stores, long before a hook reads it
                    axpy8(pre, c_tsit5_a[s + 1][s + 1], f);
                }
#pragma unroll
                for (int i = 0; i < 8; ++i) k7[i] = f[i];          // after s == 5 this is k7 = f(u_new) (FSAL)
            }
            if (twdp) { tmark = clock64(); tstage += tmark - tloop; nsteps += 1; }
            // y == u_new, k7 == f(u_new), pre == sum_{j<=6} btilde_j k_j; error estimate = h * sum btilde_j k_j
            float part = 0.f;
            {
#pragma unroll
                for (int i = 0; i < 8; ++i) part = err_term(part, pre[i], k7[i], hs, u[i], y[i], P.abstol, P.reltol);
            }
            pb ^= 1;
            S.part[pb][c][q] = part;
            {
                float m = fabsf(y[0]);
#pragma unroll
                for (int i = 1; i < 8; ++i) m = fmaxf(m, fabsf(y[i]));
                S.rng[pb][c][q] = m;
            }
            // dense output: computed / fetched by the whole warp (tensor-memory loads are warp-collective), consumed below under
            // lane-divergent control flow
            float c2[8], c3[8], c4[8];
            if (SP::SPARE_TMEM) {
                tmem_ld8f(SP::spare(tl, 0, q), c2);
                tmem_ld8f(SP::spare(tl, 1, q), c3);
                tmem_ld8f(SP::spare(tl, 2, q), c4);
                wait_ld();
            } else {
                dense_partial(tl, q, k1, c2, c3, c4);
            }
            dense_k7(k7, c2, c3, c4);
            cta_bar();
            if (twdp) { const long long n = clock64(); tb[2] += n - tmark; tmark = n; }
            if (col >= 0) {
                const float ee = fast_sqrt(((S.part[pb][c][0] + S.part[pb][c][1]) + (S.part[pb][c][2] + S.part[pb][c][3])) * (1.f / NZ));
                bool accept;
                float dtnew;
                if (!isfinite(ee)) { status = ST_NAN; accept = false; dtnew = hs; }
                else if (P.fixed_dt > 0.f) { accept = true; dtnew = P.fixed_dt; }
                else pi_controller(ee, qold, hs, accept, dtnew);      // the RECORD pass uses the same arithmetic: mse(nfc_solve(...), truth) and the loss nfc_loss_grad reports are the same number
                if (accept && !(fmaxf(fmaxf(S.rng[pb][c][0], S.rng[pb][c][1]), fmaxf(S.rng[pb][c][2], S.rng[pb][c][3])) <= SP::INPUT_BOUND)) {
                    status = ST_RANGE; accept = false;           // an accepted state outside the operand range of the split: report, never guess (the host calls re-solve it on the BF16x3 split)
                }
                if (twdp) { const long long n = clock64(); tb[4] += n - tmark; }
                if (accept) {
                    const float tn = last ? P.tf : t + hs;
                    if (RECORD) {
                        if (nacc < tape_capacity(P, col)) {
                            float* e = P.tape + (tape_row0(P, col) + nacc) * TAPE_STRIDE + 8 * q;
#pragma unroll
                            for (int hlf = 0; hlf < 2; ++hlf) {
                                const int i = 4 * hlf;
                                *reinterpret_cast<float4*>(e + i) = make_float4(u[i], u[i + 1], u[i + 2], u[i + 3]);
                                *reinterpret_cast<float4*>(e + NZ + i) = make_float4(k1[i], k1[i + 1], k1[i + 2], k1[i + 3]);
                                *reinterpret_cast<float4*>(e + 2 * NZ + i) = make_float4(c2[i], c2[i + 1], c2[i + 2], c2[i + 3]);
                                *reinterpret_cast<float4*>(e + 3 * NZ + i) = make_float4(c3[i], c3[i + 1], c3[i + 2], c3[i + 3]);
                                *reinterpret_cast<float4*>(e + 4 * NZ + i) = make_float4(c4[i], c4[i + 1], c4[i + 2], c4[i + 3]);
                            }
                            if (q == 0) { e[TAPE_T] = t; e[TAPE_H] = hs; }
                        } else {
                            status = ST_TAPE_OVERFLOW;
                        }
                    }
                    while (si < P.n_save) {
                        const float ts = sv[si];
                        if (ts > tn) break;
                        const float th = fminf(fast_div(ts - t, hs), 1.0f);
                        const long long o2i = ((long long)col * P.n_save + si) * NZ + 8 * q;
                        float v[8];
                        dense_eval(v, hs, th, u, k1, c2, c3, c4);
                        if (RECORD) {
                            const float4 ta = *reinterpret_cast<const float4*>(P.Ttrue + o2i), tb = *reinterpret_cast<const float4*>(P.Ttrue + o2i + 4);
                            const float r[8] = {v[0] - ta.x, v[1] - ta.y, v[2] - ta.z, v[3] - ta.w, v[4] - tb.x, v[5] - tb.y, v[6] - tb.z, v[7] - tb.w};
                            *reinterpret_cast<float4*>(P.resid + o2i) = make_float4(r[0], r[1], r[2], r[3]);
                            *reinterpret_cast<float4*>(P.resid + o2i + 4) = make_float4(r[4], r[5], r[6], r[7]);
                            float sse = 0.f;
#pragma unroll
                            for (int i = 0; i < 8; ++i) sse = fmaf(r[i], r[i], sse);
                            lsum += (double)sse;
                        } else if (P.Tout) {
                            float* o2 = P.Tout + o2i;
                            __stcs(reinterpret_cast<float4*>(o2), make_float4(v[0], v[1], v[2], v[3]));
                            __stcs(reinterpret_cast<float4*>(o2 + 4), make_float4(v[4], v[5], v[6], v[7]));
                        }
                        ++si;
                    }
                    if (twdp) { const long long n = clock64(); tb[5] += n - tmark; }
#pragma unroll
                    for (int i = 0; i < 8; ++i) { u[i] = y[i]; k1[i] = k7[i]; }
                    t = tn;
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
            }
        }
        if (dbgbuf && blockIdx.x == 0 && threadIdx.x == 0) {
            long long* dbg = dbgbuf;
            dbg[16] = twd[0]; dbg[17] = twd[1]; dbg[18] = twd[2]; dbg[19] = clock64() - tstart; dbg[20] = tstage; dbg[21] = nsteps; dbg[22] = tb[0]; dbg[23] = tb[1]; dbg[24] = tb[2]; dbg[25] = tb[3]; dbg[26] = tb[4]; dbg[27] = tb[5];
        }
        if (RECORD) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) lsum += __shfl_xor_sync(FULL, lsum, o);
            if (lane == 0) atomicAdd(P.loss_sum, lsum);
        }
        // release the MMA warp
        cta_bar();
        if (threadIdx.x == 0) S.stop = 1;
        cta_bar();
        warp_arrive(&S.bar_a);
    }
    fence_before_sync();
    __syncthreads();
    if (warp == NCT / 32) tmem_dealloc(S.tmem_base, TM_COLS);
}

}  // namespace tc
}  // namespace nfc
