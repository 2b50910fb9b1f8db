// nfc_embedded.cuh -- SURVEY.md 8 row f1: the "embedded" path of src/oceananigans_nn.jl:158-282, batched over columns.
//
// Per fixed step dt (reference default 600 s) and column, in PHYSICAL units:
//   1. neural-network forcing (oceananigans_nn.jl:217-223): wT = [0; inv(wT_scaling)(NN(T_scaling(T))); heat_flux],
//      T_k <- T_k - dt (wT_{k+1} - wT_k) / dz.  Oceananigans' own time stepper / closure defaults are un-vendored
//      [UPSTREAM]; forward Euler stands in for them (stated in DESIGN.md, parity unpinned).
//   2. implicit convective adjustment, convective_adjustment! (oceananigans_nn.jl:3-30) restated exactly:
//      kappa_i = K if (dT/dz at cell centre i) < 0 else 0, a = dt/dz^2,
//      ld_i = -a kappa_i (i = 2..N), ud_i = -a kappa_{i+1} (i = 1..N-1), d_i = 1 + a (kappa_i + kappa_{i+1}) (i < N), d_N = 1 + a kappa_N,
//      T <- Tridiagonal(ld, d, ud) \ T.
//      The centre gradient is the mean of the two face gradients; the bottom face carries the gradient boundary condition
//      (base_model, oceananigans_nn.jl:84), the top face of a flux boundary condition is filled with zero gradient [UPSTREAM].
// The state is carried in SCALED units x = (T - mu_T)/sigma_T (the update is affine -- the reference's matrix preserves constants except
// in its first row, which gets the corresponding affine term below -- and the sign test is scale invariant), so
// Float32 resolves the O(0.01 K) changes of a 20 C profile; temperatures are un-scaled on output only.
// One warp = 8 columns, lane == level; the 32-unknown tridiagonal systems are solved by parallel cyclic reduction in
// registers (5 rounds of shuffles) -- no sequential Thomas sweep across lanes.
#pragma once
#include "nfc_forward.cuh"

namespace nfc {

struct EmbeddedParams {
    const float* wpack;
    const float* T0;          // [B][32] physical temperature
    const float* heat_flux;   // [B]     surface temperature flux Q (K m/s), top face
    float* Tout;              // [B][n_out][32]
    long long B;
    int n_steps, save_every, n_out;
    float mu_T, sigma_T, mu_wT, sigma_wT;
    float dz, dt, K, bottom_gradient;
};

// x_i from  a_i x_{i-1} + b_i x_i + c_i x_{i+1} = d_i  (a_0 = c_31 = 0), one unknown per lane
__device__ __forceinline__ float pcr32(float a, float b, float c, float d, int lane) {
#pragma unroll
    for (int s = 1; s < NZ; s <<= 1) {
        float am = __shfl_up_sync(FULL, a, s), bm = __shfl_up_sync(FULL, b, s), cm = __shfl_up_sync(FULL, c, s), dm = __shfl_up_sync(FULL, d, s);
        float ap = __shfl_down_sync(FULL, a, s), bp = __shfl_down_sync(FULL, b, s), cp = __shfl_down_sync(FULL, c, s), dp = __shfl_down_sync(FULL, d, s);
        if (lane < s) { am = 0.f; bm = 1.f; cm = 0.f; dm = 0.f; }
        if (lane + s >= NZ) { ap = 0.f; bp = 1.f; cp = 0.f; dp = 0.f; }
        // fast division (reciprocal + multiply, 2 ulp): the diagonals are 1 + a (kappa_i + kappa_{i+1}) >= 1, far from the range limits,
        // and the off-diagonals are zero wherever the column is stable -- the IEEE sequence took its slow path for every zero numerator
        // and the solve cost three times the Chain evaluation
        const float al = -__fdividef(a, bm), ga = -__fdividef(c, bp);
        b = b + al * cm + ga * ap;
        d = d + al * dm + ga * dp;
        a = al * am;
        c = ga * cp;
    }
    return __fdividef(d, b);
}

template <class N, bool WS, int NW>
__global__ void __launch_bounds__(NW * 32, 1) embedded_kernel(const EmbeddedParams P) {
    extern __shared__ __align__(16) float smem[];
    __shared__ uint64_t bar;
    const float* sW = WS ? smem : P.wpack;
    if (WS) stage_weights(smem, P.wpack, N::P_TOTAL, &bar);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float* ws = smem + (WS ? N::P_TOTAL : 0) + warp * N::S_TOTAL;
    const long long ngroups = (P.B + CPW - 1) / CPW;
    const float inv_sT = 1.f / P.sigma_T, a = P.dt / (P.dz * P.dz), inv_dz = 1.f / P.dz;
    const float fscale = P.dt * inv_dz * inv_sT, gbot = P.bottom_gradient * inv_sT * P.dz;   // flux -> scaled increment; bottom face difference
    for (long long g = (long long)blockIdx.x * NW + warp; g < ngroups; g += (long long)gridDim.x * NW) {
        float T[CPW], Q[CPW];
#pragma unroll
        for (int c = 0; c < CPW; ++c) {
            const long long col = g * CPW + c;
            const float t0 = col < P.B ? P.T0[col * NZ + lane] : P.mu_T;
            T[c] = (t0 - P.mu_T) * inv_sT;
            Q[c] = col < P.B ? P.heat_flux[col] : 0.f;
            if (col < P.B) P.Tout[(col * P.n_out) * NZ + lane] = t0;
        }
        for (int step = 1; step <= P.n_steps; ++step) {
            nn_forward_tile<N>(sW, ws, T, lane);
            const float* sY = ws + N::S_Y;
#pragma unroll
            for (int c = 0; c < CPW; ++c) {
                // faces: F_0 = 0, F_k = sigma_wT NN[k-1] + mu_wT, F_32 = Q
                const float Flo = lane > 0 ? fmaf(P.sigma_wT, sY[c * YS + lane - 1], P.mu_wT) : 0.f;
                float Fhi = __shfl_down_sync(FULL, Flo, 1);
                if (lane == NZ - 1) Fhi = Q[c];
                float t = T[c] - fscale * (Fhi - Flo);
                // centre gradient = mean of the face gradients (bottom face: gradient BC, top face: zero gradient); only its sign matters
                const float tlo = __shfl_up_sync(FULL, t, 1), thi = __shfl_down_sync(FULL, t, 1);
                const float glo = lane > 0 ? (t - tlo) : gbot;
                const float ghi = lane < NZ - 1 ? (thi - t) : 0.f;
                const float kap = (glo + ghi < 0.f) ? P.K : 0.f;
                float kap_hi = __shfl_down_sync(FULL, kap, 1);
                if (lane == NZ - 1) kap_hi = 0.f;
                const float ld = lane > 0 ? -a * kap : 0.f;
                const float ud = lane < NZ - 1 ? -a * kap_hi : 0.f;
                const float dd = 1.f + a * (kap + kap_hi);
                // first row: d_1 keeps kappa_1 although there is no sub-diagonal (src/oceananigans_nn.jl:19-22), row sum 1 + a kappa_1: in scaled
                // units x = (T - mu)/sigma the reference's solve reads  L x' = x - (mu/sigma) a kappa_1 e_1
                if (lane == 0) t -= P.mu_T * inv_sT * a * kap;
                T[c] = pcr32(ld, dd, ud, t, lane);
            }
            __syncwarp();
            if (step % P.save_every == 0) {
                const int so = step / P.save_every;
#pragma unroll
                for (int c = 0; c < CPW; ++c) {
                    const long long col = g * CPW + c;
                    if (col < P.B) P.Tout[(col * P.n_out + so) * NZ + lane] = fmaf(P.sigma_T, T[c], P.mu_T);
                }
            }
        }
    }
}

}  // namespace nfc
