// nfc_embedded_tc.cuh -- the embedded fixed-step path (nfc_embedded.cuh) with the Chain on the tensor cores.
//
// One CTA owns a tile of 128 columns for all n_steps.  The scaled state lives in shared memory, one padded row per column; per step
//   (A) all 512 column threads (nfc_tc.cuh layout: thread = column c, quarter q) evaluate the Chain of the tile -- tc::SplitFP16x2 /
//       SplitBF16x3 ::eval, the tcgen05 MMAs issued by the MMA warp, exactly as in rhs_tc_kernel -- and leave the network output in a
//       second [128][33] array;
//   (B) the quarter-0 thread of every column (4 warps, one per SM sub-partition) does the explicit flux step and the implicit
//       convective adjustment of its column with the Thomas algorithm on 32 unknowns in registers.  (A first version transposed to
//       the SIMT kernel's lane = level layout and ran its parallel cyclic reduction: 19 K warp-instructions per tile and step against
//       1.8 K for the sequential sweeps, and as long as the Chain evaluation itself.)
// Coefficients follow convective_adjustment! (src/oceananigans_nn.jl:3-30) as in nfc_embedded.cuh; oracle: nde_oracle.py:embedded_solve.
#pragma once
#include "nfc_embedded.cuh"
#include "nfc_tc.cuh"

namespace nfc {
namespace tc {

constexpr int EMB_ROW = NZ + 1;                                      // padded row: both roles access their rows conflict-free
template <class SP>
constexpr int emb_tc_smem_bytes() { return ((SP::IMG + 127) / 128) * 128 + 2 * TILE * EMB_ROW * (int)sizeof(float); }

template <class SP>
__global__ void __launch_bounds__(NTHREADS, 1) embedded_tc_kernel(const unsigned char* __restrict__ gimg, const EmbeddedParams P) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ TcShared S;
    float* sT = reinterpret_cast<float*>(smem_raw + ((SP::IMG + 127) / 128) * 128);     // [128][33] scaled temperatures (role B -> role A)
    float* sN = sT + TILE * EMB_ROW;                                                    // [128][33] network output      (role A -> role B)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    tc_prologue(&S, smem_raw, gimg, warp, SP::IMG);
    const uint32_t smem_img = smem_u32(smem_raw);
    const float* sbias = reinterpret_cast<const float*>(smem_raw + SP::BIAS_OFF);
    const long long ntiles = (P.B + TILE - 1) / TILE;
    if (warp >= NCT / 32) {
        reg_dec<REGS_MMA>();
        if (warp == NCT / 32) mma_warp_loop<SP>(&S, smem_img, lane);
    } else {
        reg_inc<REGS_COL>();
        const int wq = warp & 3, q = warp >> 2, cA = 32 * wq + lane;       // role A: column cA, levels 8q .. 8q+7
        const float inv_sT = 1.f / P.sigma_T, a = P.dt / (P.dz * P.dz), inv_dz = 1.f / P.dz;
        const float fscale = P.dt * inv_dz * inv_sT, gbot = P.bottom_gradient * inv_sT * P.dz;
        int pd = 0;
        for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            const long long col = tile * TILE + cA;
            const bool own = q == 0, live = own && col < P.B;      // role B: the quarter-0 thread owns column cA
            float Q = 0.f;
            cta_bar();                                             // the previous tile's last reads of sT are done
            if (own) {
                float* row = sT + cA * EMB_ROW;
#pragma unroll
                for (int k = 0; k < NZ; ++k) {
                    const float t0 = live ? P.T0[col * NZ + k] : P.mu_T;
                    row[k] = (t0 - P.mu_T) * inv_sT;
                    if (live) P.Tout[(col * P.n_out) * NZ + k] = t0;
                }
                Q = live ? P.heat_flux[col] : 0.f;
            }
            for (int step = 1; step <= P.n_steps; ++step) {
                cta_bar();
                {   // (A) Chain evaluation of the tile
                    float y[8], nn[9];
#pragma unroll
                    for (int i = 0; i < 8; ++i) y[i] = sT[cA * EMB_ROW + 8 * q + i];
                    SP::eval(&S, sbias, y, nn, pd, wq, q, nullptr, false);
#pragma unroll
                    for (int i = 1; i <= 8; ++i) sN[cA * EMB_ROW + 8 * q + i - 1] = nn[i];     // neuron n = 8q + i - 1 (n = 31 is not used)
                }
                cta_bar();
                if (own) {   // (B) explicit flux step + implicit convective adjustment of column cA
                    float* row = sT + cA * EMB_ROW;
                    const float* sY = sN + cA * EMB_ROW;
                    float t[NZ], cp[NZ];
                    // faces: F_0 = 0, F_k = sigma_wT NN[k-1] + mu_wT, F_32 = Q
                    float Flo = 0.f;
#pragma unroll
                    for (int k = 0; k < NZ; ++k) {
                        const float Fhi = k < NZ - 1 ? fmaf(P.sigma_wT, sY[k], P.mu_wT) : Q;
                        t[k] = row[k] - fscale * (Fhi - Flo);
                        Flo = Fhi;
                    }
                    // kappa_k = K where the centre gradient (mean of the face gradients; bottom: gradient BC, top: zero) is negative
                    float kap_lo = ((gbot) + (t[1] - t[0]) < 0.f) ? P.K : 0.f;      // kappa_0
                    // The reference's first row has no sub-diagonal but keeps kappa_1 on its diagonal (d_1 = 1 + a (kappa_1 + kappa_2),
                    // src/oceananigans_nn.jl:19-22): its row sum is 1 + a kappa_1, so the operator does not preserve constants when the deepest
                    // cell is unstable, and the solve in scaled units x = (T - mu)/sigma needs the affine term  L x' = x - (mu/sigma) a kappa_1 e_1
                    const float aff = P.mu_T * inv_sT * a * kap_lo;                 // right-hand side of row 0 only (the sign tests use the state itself)
                    float dprev = 0.f, cprev = 0.f;
#pragma unroll
                    for (int k = 0; k < NZ; ++k) {
                        float kap_hi = 0.f;                                          // kappa_{k+1}
                        if (k < NZ - 1) {
                            const float glo = t[k + 1] - t[k];
                            const float ghi = k + 1 < NZ - 1 ? (t[k + 2] - t[k + 1]) : 0.f;
                            kap_hi = (glo + ghi < 0.f) ? P.K : 0.f;
                        }
                        const float ld = k > 0 ? -a * kap_lo : 0.f;
                        const float ud = k < NZ - 1 ? -a * kap_hi : 0.f;
                        const float dd = 1.f + a * (kap_lo + kap_hi);
                        // Thomas forward sweep (dd >= 1, off-diagonals <= 0: diagonally dominant, no pivoting)
                        const float m = fmaf(-ld, cprev, dd);
                        const float im = __fdividef(1.f, m);
                        cprev = ud * im;
                        dprev = fmaf(-ld, dprev, k == 0 ? t[k] - aff : t[k]) * im;
                        cp[k] = cprev; t[k] = dprev;
                        kap_lo = kap_hi;
                    }
#pragma unroll
                    for (int k = NZ - 2; k >= 0; --k) t[k] = fmaf(-cp[k], t[k + 1], t[k]);
#pragma unroll
                    for (int k = 0; k < NZ; ++k) row[k] = t[k];
                    if (live && step % P.save_every == 0) {
                        float* o = P.Tout + (col * P.n_out + step / P.save_every) * NZ;
#pragma unroll
                        for (int k = 0; k < NZ; k += 4)
                            *reinterpret_cast<float4*>(o + k) = make_float4(fmaf(P.sigma_T, t[k], P.mu_T), fmaf(P.sigma_T, t[k + 1], P.mu_T),
                                                                            fmaf(P.sigma_T, t[k + 2], P.mu_T), fmaf(P.sigma_T, t[k + 3], P.mu_T));
                    }
                }
            }
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
