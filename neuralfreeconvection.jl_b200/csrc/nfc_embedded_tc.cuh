// nfc_embedded_tc.cuh -- the embedded fixed-step path (nfc_embedded.cuh) with the Chain on the tensor cores.
//
// One CTA owns a tile of 128 columns for all n_steps.  Per step the 512 column threads play two roles:
//   role A (nfc_tc.cuh layout: thread = column c, quarter q): the Chain evaluation of the tile -- tc::SplitFP16x2 / SplitBF16x3 ::eval,
//           the tcgen05 MMAs issued by the MMA warp, exactly as in rhs_tc_kernel;
//   role B (nfc_embedded.cuh layout: warp w owns columns 8w..8w+7, lane = level): the explicit flux step and the implicit
//           convective adjustment -- the per-level code and the 32-unknown parallel-cyclic-reduction solve of embedded_kernel, verbatim.
// The state lives in role B's registers; it crosses to role A (scaled temperatures) and back (network output) through two
// [128][33] shared-memory arrays, one CTA barrier each way.  Everything except the network output (FP32-faithful to 1e-7) is
// therefore bit-identical to the FP32 SIMT kernel.
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
            // role B state: columns tile * 128 + 8 warp + c, lane = level
            float T[CPW], Q[CPW];
#pragma unroll
            for (int c = 0; c < CPW; ++c) {
                const long long col = tile * TILE + 8 * warp + c;
                const float t0 = col < P.B ? P.T0[col * NZ + lane] : P.mu_T;
                T[c] = (t0 - P.mu_T) * inv_sT;
                Q[c] = col < P.B ? P.heat_flux[col] : 0.f;
                if (col < P.B) P.Tout[(col * P.n_out) * NZ + lane] = t0;
            }
            for (int step = 1; step <= P.n_steps; ++step) {
#pragma unroll
                for (int c = 0; c < CPW; ++c) sT[(8 * warp + c) * EMB_ROW + lane] = T[c];
                cta_bar();
                {   // role A: Chain evaluation of the tile
                    float y[8], nn[9];
#pragma unroll
                    for (int i = 0; i < 8; ++i) y[i] = sT[cA * EMB_ROW + 8 * q + i];
                    SP::eval(&S, sbias, y, nn, pd, wq, q, nullptr, false);
#pragma unroll
                    for (int i = 1; i <= 8; ++i) sN[cA * EMB_ROW + 8 * q + i - 1] = nn[i];     // neuron n = 8q + i - 1 (n = 31 is not used)
                }
                cta_bar();
#pragma unroll
                for (int c = 0; c < CPW; ++c) {
                    const float* sY = sN + (8 * warp + c) * EMB_ROW;
                    // faces: F_0 = 0, F_k = sigma_wT NN[k-1] + mu_wT, F_32 = Q   (nfc_embedded.cuh, same operations)
                    const float Flo = lane > 0 ? fmaf(P.sigma_wT, sY[lane - 1], P.mu_wT) : 0.f;
                    float Fhi = __shfl_down_sync(FULL, Flo, 1);
                    if (lane == NZ - 1) Fhi = Q[c];
                    float t = T[c] - fscale * (Fhi - Flo);
                    const float tlo = __shfl_up_sync(FULL, t, 1), thi = __shfl_down_sync(FULL, t, 1);
                    const float glo = lane > 0 ? (t - tlo) : gbot;
                    const float ghi = lane < NZ - 1 ? (thi - t) : 0.f;
                    const float kap = (glo + ghi < 0.f) ? P.K : 0.f;
                    float kap_hi = __shfl_down_sync(FULL, kap, 1);
                    if (lane == NZ - 1) kap_hi = 0.f;
                    const float ld = lane > 0 ? -a * kap : 0.f;
                    const float ud = lane < NZ - 1 ? -a * kap_hi : 0.f;
                    const float dd = 1.f + a * (kap + kap_hi);
                    T[c] = pcr32(ld, dd, ud, t, lane);
                }
                if (step % P.save_every == 0) {
                    const int so = step / P.save_every;
#pragma unroll
                    for (int c = 0; c < CPW; ++c) {
                        const long long col = tile * TILE + 8 * warp + c;
                        if (col < P.B) P.Tout[(col * P.n_out + so) * NZ + lane] = fmaf(P.sigma_T, T[c], P.mu_T);
                    }
                }
            }
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

}  // namespace tc
}  // namespace nfc
