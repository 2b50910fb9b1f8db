// nfc_rkc_tc.cuh -- the RKC2 solve (nfc_rkc.cuh) on the tensor-core Chain of nfc_tc.cuh.
//
// Same CTA as solve_tc_kernel: 128 column slots, four threads per column (quarter q owns levels [8q, 8q+8)), the Dense layers
// as tcgen05 TS-form MMAs issued by the MMA warp, TMEM lane = column.  Every column is the state machine of nfc_rkc.cuh and a
// "tick" is one Chain evaluation of the whole tile, whatever each column uses it for (power iteration, first-step probe, RKC
// stage).  Per tick: (1) evaluation, (2) per-quarter partial reductions -> shared memory, CTA barrier, (3) the quarter-0 thread of
// every column runs rkc::transition (the same function as the FP32 kernel) on shared-memory scalar state, retires / refills,
// CTA barrier, (4) all four quarter threads apply the vector action.  y_n, f_n and the evaluation point x stay in registers;
// the power-iteration vector v and Y_{j-2} live in tensor memory (the region the Tsit5 kernel uses for k2..k7) and are loaded /
// stored by whole warps every tick -- tcgen05.ld/st are warp-collective and must never sit under a lane-dependent branch.
#pragma once
#include "nfc_rkc.cuh"
#include "nfc_tc.cuh"

namespace nfc {
namespace tc {

struct RkcShared {
    rkc::SlotsT<TILE> sl;
    float part[2][TILE][4][4];   // per quarter: r0, r1, r2, max|x| (double buffered)
    int newcol[TILE];
};

template <class SP>
constexpr int rkc_tc_smem_bytes() { return ((SP::IMG + 127) / 128) * 128 + (int)sizeof(RkcShared); }

template <class SP, bool RECORD = false>
__global__ void __launch_bounds__(NTHREADS, 1) solve_rkc_tc_kernel(const unsigned char* __restrict__ gimg, const rkc::Params P) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ TcShared S;
    RkcShared& R = *reinterpret_cast<RkcShared*>(smem_raw + ((SP::IMG + 127) / 128) * 128);   // dynamic: past the 48 KB static limit
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    tc_prologue(&S, smem_raw, gimg, warp, SP::IMG);
    const uint32_t smem_img = smem_u32(smem_raw);
    const float* sbias = reinterpret_cast<const float*>(smem_raw + SP::BIAS_OFF);
    if (warp >= NCT / 32) {
        reg_dec<REGS_MMA>();
        if (warp == NCT / 32) mma_warp_loop<SP>(&S, smem_img, lane);
    } else {
        reg_inc<REGS_COL>();
        using namespace rkc;
        const int wq = warp & 3, q = warp >> 2, c = 32 * wq + lane;
        const uint32_t tl = S.tmem_base + ((uint32_t)(32 * wq) << 16);
        const uint32_t tv = tl + TM_K + 8 * q, ty = tl + TM_K + 32 + 8 * q;
        const float reltol = P.S.reltol, abstol = P.S.abstol;
        auto* sl = &R.sl;
        int pd = 0, xb = 0, pb = 0;
        float yn[8], fn[8], x[8];
        double lsum = 0.0;          // RECORD: this thread's sum of squared residuals
        const float zero8[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};

        // scalar part of a refill (quarter-0 thread): next column from the queue, fresh controller state
        auto fill_scalar = [&]() {
            const long long nx = (long long)atomicAdd(&P.S.counters[CTR_QUEUE], 1ULL);
            if (nx < P.S.B) {
                int si = 0;
                while (si < P.S.n_save && P.S.saveat[si] <= 0.f) ++si;
                sl->col[c] = (int)nx; sl->phase[c] = PH_F0; sl->act[c] = ACT_REFILL; sl->m[c] = 0; sl->j[c] = 0; sl->it[c] = 0; sl->nstsig[c] = 0;
                sl->nacc[c] = 0; sl->nrej[c] = 0; sl->nfe[c] = 0; sl->si[c] = si; sl->status[c] = ST_OK;
                sl->newspc[c] = 1; sl->jacatt[c] = 0; sl->first[c] = 1; sl->last[c] = 0;
                sl->t[c] = 0.f; sl->absh[c] = P.S.tf; sl->h[c] = 0.f; sl->hlast[c] = 0.f; sl->errold[c] = 0.f; sl->sprad[c] = 0.f;
                sl->bot[c] = P.S.colp[nx * 3 + 0]; sl->top[c] = P.S.colp[nx * 3 + 1];
                sl->kca[c] = P.S.use_kca ? P.S.colp[nx * 3 + 2] * (float)NZ : 0.f;
            } else {
                sl->col[c] = -1; sl->act[c] = ACT_REFILL; sl->phase[c] = PH_F0; sl->bot[c] = 0.f; sl->top[c] = 0.f; sl->kca[c] = 0.f;
            }
        };
        // vector part of a refill (all four quarter threads)
        auto fill_vector = [&]() {
            const int nc = sl->col[c];
#pragma unroll
            for (int i = 0; i < 8; ++i) { yn[i] = 0.f; fn[i] = 0.f; x[i] = 0.f; }
            if (nc >= 0) {
                const float* src = P.S.T0 + (long long)nc * NZ + 8 * q;
                const float4 a = *reinterpret_cast<const float4*>(src), b = *reinterpret_cast<const float4*>(src + 4);
                yn[0] = a.x; yn[1] = a.y; yn[2] = a.z; yn[3] = a.w; yn[4] = b.x; yn[5] = b.y; yn[6] = b.z; yn[7] = b.w;
#pragma unroll
                for (int i = 0; i < 8; ++i) x[i] = yn[i];
                for (int si = 0; si < P.S.n_save && P.S.saveat[si] <= 0.f; ++si) {
                    const long long oi = ((long long)nc * P.S.n_save + si) * NZ + 8 * q;
                    if (RECORD) {
                        const float4 ta = *reinterpret_cast<const float4*>(P.S.Ttrue + oi), tb = *reinterpret_cast<const float4*>(P.S.Ttrue + oi + 4);
                        const float r[8] = {a.x - ta.x, a.y - ta.y, a.z - ta.z, a.w - ta.w, b.x - tb.x, b.y - tb.y, b.z - tb.z, b.w - tb.w};
                        *reinterpret_cast<float4*>(P.S.resid + oi) = make_float4(r[0], r[1], r[2], r[3]);
                        *reinterpret_cast<float4*>(P.S.resid + oi + 4) = make_float4(r[4], r[5], r[6], r[7]);
                        float sse = 0.f;
#pragma unroll
                        for (int i = 0; i < 8; ++i) sse = fmaf(r[i], r[i], sse);
                        lsum += (double)sse;
                    } else if (P.S.Tout) {
                        float* o = P.S.Tout + oi;
                        *reinterpret_cast<float4*>(o) = a;
                        *reinterpret_cast<float4*>(o + 4) = b;
                    }
                }
            }
        };

        if (q == 0) fill_scalar();
        tmem_st8f(tv, zero8);
        tmem_st8f(ty, zero8);
        wait_st();
        if (threadIdx.x == 0) S.any = 0;
        cta_bar();
        fill_vector();
        if (q == 0 && sl->col[c] >= 0) S.any = 1;
        cta_bar();
        while (S.any) {
            // ---------------- (1) one Chain evaluation of the tile ----------------
            xb ^= 1;
            S.xch[xb][c][q][0] = x[0];
            S.xch[xb][c][q][1] = x[7];
            float nn[9], f[8];
            SP::eval(&S, sbias, x, nn, pd, wq, q, nullptr, false);
            {
                const float ylo = q > 0 ? S.xch[xb][c][q - 1][1] : 0.f, yhi = q < 3 ? S.xch[xb][c][q + 1][0] : 0.f;
                faces_to_f(x, ylo, yhi, nn, sl->bot[c], sl->top[c], sl->kca[c], P.S.pref, q, f);
            }
            float v[8], yj[8];
            tmem_ld8f(tv, v);
            tmem_ld8f(ty, yj);
            wait_ld();
            // ---------------- (2) partial reductions of this quarter ----------------
            const int colc = sl->col[c];
            {
                float p0 = 0.f, p1 = 0.f, p2 = 0.f, p3 = 0.f;
                if (colc >= 0) {
                    const int ph = sl->phase[c];
                    if (ph == PH_F0) {
#pragma unroll
                        for (int i = 0; i < 8; ++i) { p1 = fmaf(x[i], x[i], p1); p2 = fmaf(f[i], f[i], p2); }
                    } else if (ph == PH_RHO) {
#pragma unroll
                        for (int i = 0; i < 8; ++i) { const float d = f[i] - fn[i]; p0 = fmaf(d, d, p0); }
                    } else if (ph == PH_DT0) {
#pragma unroll
                        for (int i = 0; i < 8; ++i) { const float d = (f[i] - fn[i]) / (abstol + reltol * fabsf(yn[i])); p0 = fmaf(d, d, p0); }
                    } else if (sl->j[c] >= sl->m[c]) {
                        const float h4 = 0.4f * sl->h[c];
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const float est = 0.8f * (yn[i] - x[i]) + h4 * (fn[i] + f[i]);
                            const float d = est / (abstol + reltol * fmaxf(fabsf(x[i]), fabsf(yn[i])));
                            p0 = fmaf(d, d, p0); p1 = fmaf(x[i], x[i], p1); p2 = fmaf(v[i], v[i], p2); p3 = fmaxf(p3, fabsf(x[i]));
                        }
                    }
                }
                pb ^= 1;
                *reinterpret_cast<float4*>(&R.part[pb][c][q][0]) = make_float4(p0, p1, p2, p3);
            }
            if (threadIdx.x == 0) S.any = 0;
            cta_bar();
            // ---------------- (3) scalar transition, retire / refill (quarter 0 of every column) ----------------
            if (q == 0) {
                if (colc >= 0) {
                    const float4 a = *reinterpret_cast<const float4*>(&R.part[pb][c][0][0]), b = *reinterpret_cast<const float4*>(&R.part[pb][c][1][0]);
                    const float4 d = *reinterpret_cast<const float4*>(&R.part[pb][c][2][0]), e = *reinterpret_cast<const float4*>(&R.part[pb][c][3][0]);
                    transition(P, sl, c, (a.x + b.x) + (d.x + e.x), (a.y + b.y) + (d.y + e.y), (a.z + b.z) + (d.z + e.z),
                               fmaxf(fmaxf(a.w, b.w), fmaxf(d.w, e.w)), SP::INPUT_BOUND);
                    if ((sl->act[c] & ACT_MASK) == ACT_REFILL) {
                        if (P.S.naccept) P.S.naccept[colc] = sl->nacc[c];
                        if (P.S.nreject) P.S.nreject[colc] = sl->nrej[c];
                        if (RECORD) {
                            if (sl->nacc[c] > tape_capacity(P.S, colc) && sl->status[c] == ST_OK) sl->status[c] = ST_TAPE_OVERFLOW;
                            P.S.tape_len[colc] = min(sl->nacc[c], tape_capacity(P.S, colc));
                        }
                        if (P.S.status) P.S.status[colc] = sl->status[c];
                        if (P.nfe) P.nfe[colc] = sl->nfe[c];
                        atomicAdd(&P.S.counters[CTR_ACCEPT], (unsigned long long)sl->nacc[c]);
                        atomicAdd(&P.S.counters[CTR_REJECT], (unsigned long long)sl->nrej[c]);
                        atomicAdd(&P.S.counters[CTR_NFE], (unsigned long long)sl->nfe[c]);
                        const int keep = sl->act[c] & ~ACT_MASK;     // a retiring column may still owe its last dense output
                        fill_scalar();
                        sl->act[c] |= keep;
                    }
                } else sl->act[c] = ACT_NONE;
                if (sl->col[c] >= 0) S.any = 1;
            }
            cta_bar();
            // ---------------- (4) vector action ----------------
            {
                const int act = sl->act[c];
                if (act & AF_SETF0) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) { fn[i] = f[i]; v[i] = f[i]; }
                }
                if (act & AF_ACCEPT) {
                    const float t0 = sl->o_t0[c], h = sl->o_h[c], tn = sl->o_tn[c];
                    float c2[8], c3[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float d = (x[i] - yn[i]) / h;
                        c2[i] = 3.f * d - 2.f * fn[i] - f[i]; c3[i] = -2.f * d + fn[i] + f[i];
                    }
                    if (RECORD) {          // tape row in the Tsit5 tape's power basis: u_n, c1 = f_n, c2, c3, c4 = 0, (t_n, h_n)
                        const int n = sl->o_row[c];
                        if (n < tape_capacity(P.S, colc)) {
                            float* e = P.S.tape + (tape_row0(P.S, colc) + n) * TAPE_STRIDE + 8 * q;
#pragma unroll
                            for (int hf = 0; hf < 2; ++hf) {
                                const int i = 4 * hf;
                                *reinterpret_cast<float4*>(e + i) = make_float4(yn[i], yn[i + 1], yn[i + 2], yn[i + 3]);
                                *reinterpret_cast<float4*>(e + NZ + i) = make_float4(fn[i], fn[i + 1], fn[i + 2], fn[i + 3]);
                                *reinterpret_cast<float4*>(e + 2 * NZ + i) = make_float4(c2[i], c2[i + 1], c2[i + 2], c2[i + 3]);
                                *reinterpret_cast<float4*>(e + 3 * NZ + i) = make_float4(c3[i], c3[i + 1], c3[i + 2], c3[i + 3]);
                                *reinterpret_cast<float4*>(e + 4 * NZ + i) = make_float4(0.f, 0.f, 0.f, 0.f);
                            }
                            if (q == 0) { e[TAPE_T] = t0; e[TAPE_H] = h; }
                        }
                    }
                    for (int si = sl->o_si[c]; si < P.S.n_save; ++si) {
                        const float ts = P.S.saveat[si];
                        if (ts > tn) break;
                        const float th = fminf((ts - t0) / h, 1.0f);
                        float o[8];
#pragma unroll
                        for (int i = 0; i < 8; ++i) o[i] = fmaf(h * th, fmaf(th, fmaf(th, c3[i], c2[i]), fn[i]), yn[i]);
                        const long long oi = ((long long)colc * P.S.n_save + si) * NZ + 8 * q;
                        if (RECORD) {
                            const float4 ta = *reinterpret_cast<const float4*>(P.S.Ttrue + oi), tb = *reinterpret_cast<const float4*>(P.S.Ttrue + oi + 4);
                            const float r[8] = {o[0] - ta.x, o[1] - ta.y, o[2] - ta.z, o[3] - ta.w, o[4] - tb.x, o[5] - tb.y, o[6] - tb.z, o[7] - tb.w};
                            *reinterpret_cast<float4*>(P.S.resid + oi) = make_float4(r[0], r[1], r[2], r[3]);
                            *reinterpret_cast<float4*>(P.S.resid + oi + 4) = make_float4(r[4], r[5], r[6], r[7]);
                            float sse = 0.f;
#pragma unroll
                            for (int i = 0; i < 8; ++i) sse = fmaf(r[i], r[i], sse);
                            lsum += (double)sse;
                        } else if (P.S.Tout) {
                            float* dst = P.S.Tout + oi;
                            *reinterpret_cast<float4*>(dst) = make_float4(o[0], o[1], o[2], o[3]);
                            *reinterpret_cast<float4*>(dst + 4) = make_float4(o[4], o[5], o[6], o[7]);
                        }
                    }
#pragma unroll
                    for (int i = 0; i < 8; ++i) { yn[i] = x[i]; fn[i] = f[i]; }
                }
                if (act & AF_RHO_DONE) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) v[i] -= yn[i];
                }
                const int a = act & ACT_MASK;
                const float ca = sl->ca[c], cb = sl->cb[c], cc = sl->cc[c];
                if (a == ACT_RHO_INIT) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) { v[i] = fmaf(ca, yn[i], fmaf(cb, v[i], cc)); x[i] = v[i]; }
                } else if (a == ACT_RHO_CONT) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) { v[i] = fmaf(f[i] - fn[i], ca, yn[i]); x[i] = v[i]; }
                } else if (a == ACT_DT0) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) x[i] = fmaf(ca, fn[i], yn[i]);
                } else if (a == ACT_BEGIN) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) { yj[i] = yn[i]; x[i] = fmaf(ca, fn[i], yn[i]); }
                } else if (a == ACT_STAGE) {
                    const float cd = sl->cd[c], ce = sl->ce[c];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float y1 = ca * x[i] + cb * yj[i] + cc * yn[i] + (cd * f[i] + ce * fn[i]);
                        yj[i] = x[i]; x[i] = y1;
                    }
                } else if (a == ACT_REFILL) {
                    fill_vector();
#pragma unroll
                    for (int i = 0; i < 8; ++i) { v[i] = 0.f; yj[i] = 0.f; }
                }
            }
            tmem_st8f(tv, v);
            tmem_st8f(ty, yj);
            wait_st();
        }
        if (RECORD) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) lsum += __shfl_xor_sync(FULL, lsum, o);
            if (lane == 0) atomicAdd(P.S.loss_sum, lsum);
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
