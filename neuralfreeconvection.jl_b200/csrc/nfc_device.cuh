// nfc_device.cuh -- device-side building blocks shared by every kernel of libnfc_b200 (sm_100a only).
//
// Mapping used by all SIMT kernels ("warp tile"):
//   * one warp owns CPW = 8 ocean columns at a time;
//   * ODE work (stage combinations, flux divergence, error norm, dense output): lane == vertical level
//     (Nz = 32 == warp width), the 8 columns live in registers u[c], k_j[c];
//   * Dense layers: lane == output neuron (neurons lane + 32 j), the 8 columns are the register-tile's
//     second dimension: every weight fetched from shared memory feeds 8 FMAs, every activation 4..8.
//   * weights live in shared memory in a lane-permuted layout (one LDS.128 per lane per k), staged once
//     per CTA with a TMA bulk copy (cp.async.bulk + mbarrier).
//
// Reference semantics implemented here: src/free_convection_nde.jl:29-38 and
// src/convective_adjustment_nde.jl:33-48 (RHS), Flux Dense/Conv + destructure layout.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace nfc {

constexpr int NZ = 32;          // vertical levels == warp width
constexpr int CPW = 8;          // columns per warp tile
constexpr int YS = 33;          // row stride of the NN-output staging buffer [CPW][YS]
constexpr unsigned FULL = 0xffffffffu;

// ---- Tsit5 (Tsitouras 2011; OrdinaryDiffEq 5.60.1) ------------------------------------------------
#define NFC_A21 0.161f
#define NFC_A31 (-0.008480655492356989f)
#define NFC_A32 0.335480655492357f
#define NFC_A41 2.8971530571054935f
#define NFC_A42 (-6.359448489975075f)
#define NFC_A43 4.3622954328695815f
#define NFC_A51 5.325864828439257f
#define NFC_A52 (-11.748883564062828f)
#define NFC_A53 7.4955393428898365f
#define NFC_A54 (-0.09249506636175525f)
#define NFC_A61 5.86145544294642f
#define NFC_A62 (-12.92096931784711f)
#define NFC_A63 8.159367898576159f
#define NFC_A64 (-0.071584973281401f)
#define NFC_A65 (-0.028269050394068383f)
#define NFC_B1 0.09646076681806523f
#define NFC_B2 0.01f
#define NFC_B3 0.4798896504144996f
#define NFC_B4 1.379008574103742f
#define NFC_B5 (-3.290069515436081f)
#define NFC_B6 2.324710524099774f
#define NFC_BT1 (-0.00178001105222577714f)
#define NFC_BT2 (-0.0008164344596567469f)
#define NFC_BT3 0.007880878010261995f
#define NFC_BT4 (-0.1447110071732629f)
#define NFC_BT5 0.5823571654525552f
#define NFC_BT6 (-0.45808210592918697f)
#define NFC_BT7 0.015151515151515152f

__device__ __forceinline__ float tsit5_c(int s) {   // c_s, s = 1..7
    switch (s) { case 1: return 0.f; case 2: return 0.161f; case 3: return 0.327f; case 4: return 0.9f;
                 case 5: return 0.9800255409045097f; default: return 1.f; }
}

// a_{s,j} (s = 2..7 ; row 7 = b), compile-time so the stage loops fully unroll into immediates
template <int S, int J> struct TsA { static constexpr float v = 0.f; };
#define NFC_TSA(S, J, V) template <> struct TsA<S, J> { static constexpr float v = V; };
NFC_TSA(2, 1, NFC_A21)
NFC_TSA(3, 1, NFC_A31) NFC_TSA(3, 2, NFC_A32)
NFC_TSA(4, 1, NFC_A41) NFC_TSA(4, 2, NFC_A42) NFC_TSA(4, 3, NFC_A43)
NFC_TSA(5, 1, NFC_A51) NFC_TSA(5, 2, NFC_A52) NFC_TSA(5, 3, NFC_A53) NFC_TSA(5, 4, NFC_A54)
NFC_TSA(6, 1, NFC_A61) NFC_TSA(6, 2, NFC_A62) NFC_TSA(6, 3, NFC_A63) NFC_TSA(6, 4, NFC_A64) NFC_TSA(6, 5, NFC_A65)
NFC_TSA(7, 1, NFC_B1) NFC_TSA(7, 2, NFC_B2) NFC_TSA(7, 3, NFC_B3) NFC_TSA(7, 4, NFC_B4) NFC_TSA(7, 5, NFC_B5) NFC_TSA(7, 6, NFC_B6)
#undef NFC_TSA

// dense output weights b_j(theta), j = 1..7 (OrdinaryDiffEq Tsit5 interpolant; order conditions checked in tests)
__device__ __forceinline__ void tsit5_btheta(float th, float (&b)[7]) {
    const float t2 = th * th, t3 = t2 * th, t4 = t2 * t2;
    b[0] = th + (-2.763706197274826f) * t2 + 2.9132554618219126f * t3 + (-1.0530884977290216f) * t4;
    b[1] = 0.13169999999999998f * t2 + (-0.2234f) * t3 + 0.1017f * t4;
    b[2] = 3.9302962368947516f * t2 + (-5.941033872131505f) * t3 + 2.490627285651253f * t4;
    b[3] = (-12.411077166933676f) * t2 + 30.33818863028232f * t3 + (-16.548102889244902f) * t4;
    b[4] = 37.50931341651104f * t2 + (-88.1789048947664f) * t3 + 47.37952196281928f * t4;
    b[5] = (-27.896526289197286f) * t2 + 65.09189467479366f * t3 + (-34.87065786149661f) * t4;
    b[6] = 1.5f * t2 + (-4.0f) * t3 + 2.5f * t4;
}

// Step-boundary arithmetic shared by the adaptive kernels of the default path (one- / two-tile / small forward kernels agree with their
// RECORD instantiations; the three backward kernels with each other).  Divisions and the square root of the
// controller are the fast approximations (MUFU.RCP / MUFU.SQRT + one multiply, <= 2 ulp): the step-size sequence does not need more, the
// accept test compares the error norm itself, and the position inside a step (theta of the dense output) is good to 1e-7.  The IEEE
// sequences they replace were ~10 instructions each on the serial path every thread walks between two barriers.
__device__ __forceinline__ float fast_div(float a, float b) { return __fdividef(a, b); }
__device__ __forceinline__ float fast_sqrt(float x) { float r; asm("sqrt.approx.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
// squared scaled error of one level: (h * sum btilde_j k_j / (abstol + max(|u|, |u_new|) reltol))^2 accumulated into `part`
__device__ __forceinline__ float err_term(float part, float e6, float k7, float hs, float u, float y, float abstol, float reltol) {
    const float ei = fmaf(NFC_BT7, k7, e6) * hs;
    const float r = fast_div(ei, fmaf(fmaxf(fabsf(u), fabsf(y)), reltol, abstol));
    return fmaf(r, r, part);
}
// OrdinaryDiffEq's PI controller (beta1 = 7/50, beta2 = 2/25, gamma = 9/10, qmin = 1/5, qmax = 10; reject: dt / min(1/qmin, q11 / gamma))
__device__ __forceinline__ void pi_controller(float ee, float qold, float hs, bool& accept, float& dtnew) {
    const float q11 = __powf(ee, 0.14f);
    float qq = fast_div(q11, __powf(qold, 0.08f));
    qq = fmaxf(0.1f, fminf(5.0f, qq * (1.0f / 0.9f)));
    accept = ee <= 1.0f;
    dtnew = accept ? fast_div(hs, qq) : fast_div(hs, fminf(5.0f, q11 * (1.0f / 0.9f)));
}

// ---- network shapes ---------------------------------------------------------------------------------
// All five reference Chains (train_free_convection_nde.jl:125-153) are
//   [Conv((CONVW,1),1=>1,relu)] -> Dense(IN0,H,relu) -> (NHID-1) x Dense(H,H,relu) -> Dense(H,Nz-1)
template <int H_, int NHID_, int CONVW_>
struct Net {
    static constexpr int H = H_, NHID = NHID_, CONVW = CONVW_;
    static constexpr int IN0 = NZ - (CONVW_ ? CONVW_ - 1 : 0);
    static constexpr int NOUT = NZ - 1;
    static constexpr int NPL = H_ / 32;                 // hidden neurons per lane
    // packed (shared-memory) layout, in floats.  Rows are padded by 4 floats (RSW = H+4, RSL = 36): with a
    // row stride == 4 (mod 32) banks, both the forward access (lane -> 4 consecutive floats of ONE row) and the
    // transposed access of the adjoint (lane -> its OWN row, same 4 columns) are conflict-free LDS.128.
    static constexpr int RSW = H_ + 4;                  // row stride of [*][H] matrices
    static constexpr int RSL = 36;                      // row stride of the last layer [H][32]
    static constexpr int P_CONV = 0;                    // CONVW weights + bias, padded to 8
    static constexpr int P_W0 = 8;                      // [IN0][RSW] lane-permuted columns
    static constexpr int P_B0 = P_W0 + IN0 * RSW;       // [H] lane-permuted
    static constexpr int P_HID = P_B0 + H_;             // (NHID-1) x { [H][RSW], [H] }
    static constexpr int HIDP_STRIDE = H_ * RSW + H_;
    static constexpr int P_WL = P_HID + (NHID_ - 1) * HIDP_STRIDE;   // [H][RSL] (column 31.. zero)
    static constexpr int P_BL = P_WL + H_ * RSL;        // [32]
    static constexpr int P_TOTAL = P_BL + 32;           // multiple of 4 floats (16 B) by construction
    static constexpr int HID_STRIDE = H_ * H_ + H_;     // theta stride of one hidden layer
    // Flux.destructure layout, in floats
    static constexpr int T_CONV = 0;
    static constexpr int T_W0 = CONVW_ ? CONVW_ + 1 : 0;
    static constexpr int T_B0 = T_W0 + IN0 * H_;
    static constexpr int T_HID = T_B0 + H_;
    static constexpr int T_WL = T_HID + (NHID_ - 1) * (H_ * H_ + H_);
    static constexpr int T_BL = T_WL + H_ * NOUT;
    static constexpr int T_TOTAL = T_BL + NOUT;
    // per-warp activation scratch, in floats: X[32][8] | A[H][8] | B[H][8] | Y[8][33]
    static constexpr int S_X = 0;
    static constexpr int S_A = NZ * CPW;
    static constexpr int S_B = S_A + H_ * CPW;
    static constexpr int S_Y = S_B + H_ * CPW;
    static constexpr int S_TOTAL = ((S_Y + CPW * YS + 3) / 4) * 4;
};

// packed index -> theta index (or -1 for zero padding); used by the pack kernel
template <class N>
__host__ __device__ inline int packed_to_theta(int p) {
    if (p < N::P_W0) {
        if (N::CONVW == 0 || p > N::CONVW) return -1;
        return N::T_CONV + p;                                  // w[0..CONVW-1], bias at CONVW
    }
    if (p < N::P_B0) {                                          // W0[k][NPL*lane + j] = W[k][lane + 32 j]
        int q = p - N::P_W0, k = q / N::RSW, r = q % N::RSW;
        if (r >= N::H) return -1;
        int lane = r / N::NPL, j = r % N::NPL;
        return N::T_W0 + k * N::H + lane + 32 * j;
    }
    if (p < N::P_HID) { int r = p - N::P_B0, lane = r / N::NPL, j = r % N::NPL; return N::T_B0 + lane + 32 * j; }
    if (p < N::P_WL) {
        int q = p - N::P_HID, l = q / N::HIDP_STRIDE, r0 = q % N::HIDP_STRIDE;
        int tbase = N::T_HID + l * N::HID_STRIDE;
        if (r0 < N::H * N::RSW) {
            int k = r0 / N::RSW, r = r0 % N::RSW;
            if (r >= N::H) return -1;
            int lane = r / N::NPL, j = r % N::NPL;
            return tbase + k * N::H + lane + 32 * j;
        }
        int r = r0 - N::H * N::RSW, lane = r / N::NPL, j = r % N::NPL;
        return tbase + N::H * N::H + lane + 32 * j;
    }
    if (p < N::P_BL) { int q = p - N::P_WL, k = q / N::RSL, n = q % N::RSL; return n < N::NOUT ? N::T_WL + k * N::NOUT + n : -1; }
    int n = p - N::P_BL;
    return n < N::NOUT ? N::T_BL + n : -1;
}

// ---- TMA bulk copy global -> shared (SASS: UBLKCP) + mbarrier ----------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tNFC_WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra NFC_DONE_%=;\n\tbra NFC_WAIT_%=;\n\tNFC_DONE_%=:\n\t}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

// Stage `nfloats` packed weights into shared memory with the bulk-copy engine; all threads return after the data landed.
__device__ __forceinline__ void stage_weights(float* sW, const float* __restrict__ gW, int nfloats, uint64_t* bar) {
    if (threadIdx.x == 0) mbar_init(bar, 1);
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned total = (unsigned)nfloats * 4u;
        mbar_expect_tx(bar, total);
        constexpr unsigned CHUNK = 32768u;
        for (unsigned off = 0; off < total; off += CHUNK) {
            unsigned n = total - off < CHUNK ? total - off : CHUNK;
            bulk_g2s(reinterpret_cast<char*>(sW) + off, reinterpret_cast<const char*>(gW) + off, n, bar);
        }
    }
    mbar_wait(bar, 0);
}

// The same for a barrier that is already initialised (count 1) and has completed `parity` phases mod 2: used when a CTA
// replaces its weights (ensemble members).  The caller's __syncthreads() guarantees nobody still reads the old weights.
__device__ __forceinline__ void restage_weights(float* sW, const float* __restrict__ gW, int nfloats, uint64_t* bar, unsigned parity) {
    if (threadIdx.x == 0) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy reads of sW before the async-proxy overwrite
        const unsigned total = (unsigned)nfloats * 4u;
        mbar_expect_tx(bar, total);
        constexpr unsigned CHUNK = 32768u;
        for (unsigned off = 0; off < total; off += CHUNK) {
            unsigned n = total - off < CHUNK ? total - off : CHUNK;
            bulk_g2s(reinterpret_cast<char*>(sW) + off, reinterpret_cast<const char*>(gW) + off, n, bar);
        }
    }
    mbar_wait(bar, parity);
}

// ---- Dense layers on a warp tile ----------------------------------------------------------------------
// out[n][c] = act(b[n] + sum_k W[k][n] in[k][c]),  in: [K][8], out: [32*NPL][8] (row = neuron), weights lane-permuted.
template <int K, int NPL, int RS, bool RELU>
__device__ __forceinline__ void dense_hidden(const float* __restrict__ W, const float* __restrict__ Bv, const float* __restrict__ in,
                                             float* __restrict__ out, int lane) {
    static_assert(NPL == 4 || NPL == 8, "hidden width must be 128 or 256");
    float acc[NPL][CPW];
#pragma unroll
    for (int j4 = 0; j4 < NPL / 4; ++j4) {
        const float4 b = *reinterpret_cast<const float4*>(Bv + NPL * lane + 4 * j4);
#pragma unroll
        for (int c = 0; c < CPW; ++c) { acc[4 * j4 + 0][c] = b.x; acc[4 * j4 + 1][c] = b.y; acc[4 * j4 + 2][c] = b.z; acc[4 * j4 + 3][c] = b.w; }
    }
    const float* wp = W + NPL * lane;
#pragma unroll 4
    for (int k = 0; k < K; ++k) {
        const float4 x0 = *reinterpret_cast<const float4*>(in + k * CPW);
        const float4 x1 = *reinterpret_cast<const float4*>(in + k * CPW + 4);
        const float x[CPW] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
#pragma unroll
        for (int j4 = 0; j4 < NPL / 4; ++j4) {
            const float4 w4 = *reinterpret_cast<const float4*>(wp + k * RS + 4 * j4);
            const float w[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int c = 0; c < CPW; ++c) acc[4 * j4 + j][c] = fmaf(w[j], x[c], acc[4 * j4 + j][c]);
        }
    }
#pragma unroll
    for (int j = 0; j < NPL; ++j) {
        float4 o0, o1;
        if (RELU) {
            o0 = make_float4(fmaxf(acc[j][0], 0.f), fmaxf(acc[j][1], 0.f), fmaxf(acc[j][2], 0.f), fmaxf(acc[j][3], 0.f));
            o1 = make_float4(fmaxf(acc[j][4], 0.f), fmaxf(acc[j][5], 0.f), fmaxf(acc[j][6], 0.f), fmaxf(acc[j][7], 0.f));
        } else {
            o0 = make_float4(acc[j][0], acc[j][1], acc[j][2], acc[j][3]);
            o1 = make_float4(acc[j][4], acc[j][5], acc[j][6], acc[j][7]);
        }
        float* o = out + (lane + 32 * j) * CPW;
        *reinterpret_cast<float4*>(o) = o0;
        *reinterpret_cast<float4*>(o + 4) = o1;
    }
}

// last layer Dense(H, Nz-1) (identity): lane == output neuron (lane 31 is zero padding); result -> Y[c][lane]
template <int K, int RS>
__device__ __forceinline__ void dense_last(const float* __restrict__ W, const float* __restrict__ Bv, const float* __restrict__ in,
                                           float* __restrict__ Y, int lane) {
    float acc[CPW];
    const float b = Bv[lane];
#pragma unroll
    for (int c = 0; c < CPW; ++c) acc[c] = b;
#pragma unroll 8
    for (int k = 0; k < K; ++k) {
        const float w = W[k * RS + lane];
        const float4 x0 = *reinterpret_cast<const float4*>(in + k * CPW);
        const float4 x1 = *reinterpret_cast<const float4*>(in + k * CPW + 4);
        acc[0] = fmaf(w, x0.x, acc[0]); acc[1] = fmaf(w, x0.y, acc[1]); acc[2] = fmaf(w, x0.z, acc[2]); acc[3] = fmaf(w, x0.w, acc[3]);
        acc[4] = fmaf(w, x1.x, acc[4]); acc[5] = fmaf(w, x1.y, acc[5]); acc[6] = fmaf(w, x1.z, acc[6]); acc[7] = fmaf(w, x1.w, acc[7]);
    }
#pragma unroll
    for (int c = 0; c < CPW; ++c) Y[c * YS + lane] = acc[c];
}

// NN(T) for the warp's 8 columns.  y[c] = T value of column c at level `lane`.  Result: sY[c][n], n = 0..30.
template <class N>
__device__ __forceinline__ void nn_forward_tile(const float* __restrict__ sW, float* __restrict__ ws, const float (&y)[CPW], int lane) {
    float* sX = ws + N::S_X;
    float* sA = ws + N::S_A;
    float* sB = ws + N::S_B;
    float* sY = ws + N::S_Y;
    if (N::CONVW == 0) {
        *reinterpret_cast<float4*>(sX + lane * CPW) = make_float4(y[0], y[1], y[2], y[3]);
        *reinterpret_cast<float4*>(sX + lane * CPW + 4) = make_float4(y[4], y[5], y[6], y[7]);
    } else {
        // Flux Conv((k,1),1=>1,relu): x_i = relu(b + sum_j w[j] * T[i + (k-1-j)]), i = 0..IN0-1 (flipped kernel)
        float x[CPW];
        const float cb = sW[N::P_CONV + N::CONVW];
#pragma unroll
        for (int c = 0; c < CPW; ++c) {
            float acc = 0.f;
#pragma unroll
            for (int j = 0; j < N::CONVW; ++j) {
                const float v = __shfl_down_sync(FULL, y[c], N::CONVW - 1 - j);
                acc = fmaf(sW[N::P_CONV + j], v, acc);
            }
            x[c] = fmaxf(acc + cb, 0.f);
        }
        *reinterpret_cast<float4*>(sX + lane * CPW) = make_float4(x[0], x[1], x[2], x[3]);
        *reinterpret_cast<float4*>(sX + lane * CPW + 4) = make_float4(x[4], x[5], x[6], x[7]);
    }
    __syncwarp();
    dense_hidden<N::IN0, N::NPL, N::RSW, true>(sW + N::P_W0, sW + N::P_B0, sX, sA, lane);
    __syncwarp();
    float* in = sA;
    float* out = sB;
#pragma unroll
    for (int l = 0; l < N::NHID - 1; ++l) {
        dense_hidden<N::H, N::NPL, N::RSW, true>(sW + N::P_HID + l * N::HIDP_STRIDE, sW + N::P_HID + l * N::HIDP_STRIDE + N::H * N::RSW, in, out, lane);
        __syncwarp();
        float* t = in; in = out; out = t;
    }
    dense_last<N::H, N::RSL>(sW + N::P_WL, sW + N::P_BL, in, sY, lane);
    __syncwarp();
}

// Backward solve, accepted step: a step cut short by a save time (h below the controller's proposal dtp) says nothing about the step
// size the problem allows, so it neither enters the controller's error history nor replaces the proposal (growth capped at 10 h, the
// controller's own qmax).  Otherwise the PI controller sees "tiny error, then normal error" at every one of the 128 save times and keeps
// the steps ~ 25 % shorter than the tolerance asks for (oracle/nde_oracle.c:adjoint_continuous does the same).
__device__ __forceinline__ void adj_accept_update(float ee, float h, float dtp, float dtnew, bool adaptive, float& qold, float& dt) {
    const bool cut = adaptive && h < dtp;
    if (adaptive && !cut) qold = fmaxf(ee, 1e-4f);
    dt = cut ? fmaxf(dtnew, fminf(dtp, 10.f * h)) : dtnew;
}

// Backward solve, rejected step: lambda' = -J(u(t))^T lambda has a right-hand side that JUMPS in time wherever a temperature gradient
// crosses zero (the active set of min(0, K dT/dz)), so across such a time the error estimate falls like h, not h^5, and the controller's
// own shrink (~ 15 % per rejection at EEst ~ 2) needs 6-7 rejections in a row -- each one costs a cancel pass on top.  From the second
// rejection in a row on, the new step assumes first-order behaviour: 0.9 h / EEst (at most 5 x shorter, the controller's qmin).
__device__ __forceinline__ float adj_reject_dt(float ee, float h, float dtnew, int rejected_in_a_row) {
    return (rejected_in_a_row >= 1 && isfinite(ee)) ? fminf(dtnew, fmaxf(0.2f * h, 0.9f * h / ee)) : dtnew;
}

// Flux divergence + convective adjustment for one column (lane == level k):
//   F_lo = k==0 ? bottom : NN[k-1] - min(0, K_CA * (T_k - T_{k-1}) / dz),   f_k = -pref * (F_hi - F_lo) / dz,  1/dz = Nz
// nz < 32 (--grid-points, train_free_convection_nde.jl:29,90): the column lives in levels [0, nz) of the 32-level layout (the Chain is
// zero-padded on the host, nfc_api.cu); the top flux sits on face nz and the levels above do not move.
__device__ __forceinline__ float flux_divergence(float yk, const float* __restrict__ sYc, float bottom, float top, float kca, float pref, int lane,
                                                 int nz = NZ) {
    const float ylo = __shfl_up_sync(FULL, yk, 1);
    const float fnz = (float)nz;
    float Flo = bottom;
    if (lane > 0) {
        const float d = kca * ((yk - ylo) * fnz);
        Flo = sYc[lane - 1] - fminf(d, 0.f);
    }
    float Fhi = __shfl_down_sync(FULL, Flo, 1);
    if (lane == nz - 1) Fhi = top;
    const float f = -pref * ((Fhi - Flo) * fnz);
    return lane < nz ? f : 0.f;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}

}  // namespace nfc
