// nfc_rkc.cuh -- SURVEY row f2: a stabilised explicit integrator in the role ROCK4 plays in the reference's production runs
// (train_free_convection_nde.sh:3, train_free_convection_nde.jl:328, figureC:31).  ROCK4's coefficient tables are
// OrdinaryDiffEq's [UPSTREAM, not vendored]; this is the second-order Runge-Kutta-Chebyshev method (RKC; Sommeijer, Shampine &
// Verwer 1998) whose coefficients are closed-form Chebyshev recurrences (host: rkc_coefficients in nfc_api.cu, the same table
// as oracle/nde_oracle.py:rkc_coefficients).  CPU twin of this kernel: oracle/nde_oracle.py:rkc_solve.
//
// Same warp layout as solve_kernel (one warp = 8 column slots, lane = level for the ODE, lane = neuron for the Dense layers,
// slots refilled from a global queue), but every slot is its own little state machine, because the number of stages differs per
// column and step and because the spectral-radius power iteration and the first-step probe are RHS evaluations too:
//
//     one "tick" = one RHS evaluation of the warp's 8 slots, each on its own input x and for its own purpose (phase)
//       PH_F0     f(y_n) of a fresh column
//       PH_RHO    one round of the nonlinear power iteration  sigma = |f(v) - f(y_n)| / |v - y_n|
//       PH_DT0    first-step probe  f(y_n + h f_n)
//       PH_STAGE  stage j of an m-stage RKC step; the m-th evaluation is f(y_{n+1}) (error estimate, FSAL)
//
// After the evaluation: (B) the reductions the phase needs, (C) the scalar transition of all 8 slots side by side on lanes
// 0..7, which leaves an action code and up to five coefficients per slot, (D) the vector update of every slot.
#pragma once
#include "nfc_forward.cuh"

namespace nfc {
namespace rkc {

constexpr int MMAX = 64;                      // table size; per-solve cap = min(MMAX, sqrt(reltol / (10 eps)))
constexpr int TAB_STRIDE = 8;                 // floats per (m, j): mu, nu, 1-mu-nu, mu~, gamma~  (mu~_1 at j = 1)
constexpr int TAB_FLOATS = (MMAX + 1) * (MMAX + 1) * TAB_STRIDE;
enum : int { CTR_NFE = 6 };
enum : int { PH_F0 = 0, PH_RHO = 1, PH_DT0 = 2, PH_STAGE = 3 };
enum : int { ACT_NONE = 0, ACT_RHO_INIT = 1, ACT_RHO_CONT = 2, ACT_DT0 = 3, ACT_BEGIN = 4, ACT_STAGE = 5, ACT_REFILL = 6, ACT_MASK = 15,
             AF_ACCEPT = 16, AF_RHO_DONE = 32, AF_SETF0 = 64 };

template <int NS>
struct SlotsT {   // shared memory; index = slot (SIMT kernel: 8 per warp; tensor-core kernel: 128 per CTA)
    int col[NS], phase[NS], act[NS], m[NS], j[NS], it[NS], nstsig[NS], nacc[NS], nrej[NS], nfe[NS], si[NS], status[NS];
    int newspc[NS], jacatt[NS], first[NS], last[NS];
    float t[NS], absh[NS], h[NS], hlast[NS], errold[NS], sprad[NS], ynrm2[NS], vnrm2[NS], dynrm[NS];
    float sigma[NS], env[NS], smax[NS];
    float bot[NS], top[NS], kca[NS];
    float ca[NS], cb[NS], cc[NS], cd[NS], ce[NS];   // coefficients of the vector action
    float o_t0[NS], o_h[NS], o_tn[NS];              // the step just accepted (dense output): start, size, end
    int o_si[NS];                                   //   and the first save index it may serve
    int o_row[NS];                                  //   and its tape row (RECORD mode: accepted-step index)
};
using Slots = SlotsT<CPW>;

struct Params {
    SolveParams S;            // T0, colp, saveat, Tout, naccept, nreject, status, counters, B, n_save, max_iters, use_kca, pref, tf, tolerances
    const float* tab;         // [MMAX+1][MMAX+1][TAB_STRIDE]
    int* nfe;                 // [B] RHS evaluations per column (may be null)
    int mmax;
};

template <bool RECORD = false>
__device__ __forceinline__ void fill(const Params& P, Slots* sl, int c, float& yn, float& x, int lane, double* lsum = nullptr) {
    long long next = 0;
    if (lane == 0) next = (long long)atomicAdd(&P.S.counters[CTR_QUEUE], 1ULL);
    next = __shfl_sync(FULL, next, 0);
    if (next < P.S.B) {
        yn = P.S.T0[next * NZ + lane];
        x = yn;
        int si = 0;
        while (si < P.S.n_save && P.S.saveat[si] <= 0.f) {
            const long long o = (next * P.S.n_save + si) * NZ + lane;
            if (RECORD) { const float r = yn - P.S.Ttrue[o]; P.S.resid[o] = r; *lsum += (double)warp_sum(r * r); }
            else if (P.S.Tout) P.S.Tout[o] = yn;
            ++si;
        }
        if (lane == 0) {
            sl->col[c] = (int)next; sl->phase[c] = PH_F0; sl->act[c] = ACT_NONE; sl->m[c] = 0; sl->j[c] = 0; sl->it[c] = 0; sl->nstsig[c] = 0;
            sl->nacc[c] = 0; sl->nrej[c] = 0; sl->nfe[c] = 0; sl->si[c] = si; sl->status[c] = ST_OK;
            sl->newspc[c] = 1; sl->jacatt[c] = 0; sl->first[c] = 1; sl->last[c] = 0;
            sl->t[c] = 0.f; sl->absh[c] = P.S.tf; sl->h[c] = 0.f; sl->hlast[c] = 0.f; sl->errold[c] = 0.f; sl->sprad[c] = 0.f;
            sl->bot[c] = P.S.colp[next * 3 + 0]; sl->top[c] = P.S.colp[next * 3 + 1];
            sl->kca[c] = P.S.use_kca ? P.S.colp[next * 3 + 2] : 0.f;
        }
    } else {
        yn = 0.f; x = 0.f;
        if (lane == 0) { sl->col[c] = -1; sl->act[c] = ACT_NONE; sl->phase[c] = PH_F0; sl->bot[c] = 0.f; sl->top[c] = 0.f; sl->kca[c] = 0.f; }
    }
}

// Scalar transition of slot s after its evaluation; r0, r1, r2 are the reductions of step (B); xmax = max |y_{n+1}| is checked
// against `bound` when a step is accepted (operand range of the tensor-core splits; +inf for the FP32 kernel).
template <class SL>
__device__ __forceinline__ void transition(const Params& P, SL* sl, int s, float r0, float r1, float r2, float xmax, float bound) {
    const float tf = P.S.tf, hmax = tf, uround = 1.1920929e-07f, hmin = 10.f * uround * tf;
    const int ph = sl->phase[s];
    int act = ACT_NONE;
    float tcur = sl->t[s];
    float absh = sl->absh[s];
    bool to_rho = false, to_after_rho = false, to_begin = false, retire = false;
    sl->nfe[s] += 1;
    if (ph == PH_F0) {
        act |= AF_SETF0;
        sl->ynrm2[s] = r1; sl->vnrm2[s] = r2;
        to_rho = true;
    } else if (ph == PH_RHO) {
        const int it = sl->it[s] + 1;
        sl->it[s] = it;
        const float dynrm = sl->dynrm[s];
        const float dfnrm = sqrtf(r0);
        const float sigmal = sl->sigma[s], sigma = dfnrm / dynrm;
        const float envl = sl->env[s], env = fmaxf(sigma, sigmal);
        float smax = fmaxf(sl->smax[s], sigma);
        sl->sigma[s] = sigma; sl->env[s] = env;
        const bool conv = it >= 3 && fabsf(env - envl) <= fmaxf(env, 1.0f / hmax) * 0.01f;
        if (conv) smax = env;
        sl->smax[s] = smax;
        if (conv || it >= 50) {
            act |= AF_RHO_DONE;
            if (!isfinite(smax)) { sl->status[s] = ST_NAN; retire = true; }
            else { sl->sprad[s] = 1.2f * smax; sl->jacatt[s] = 1; to_after_rho = true; }
        } else if (dfnrm != 0.f) {
            act |= ACT_RHO_CONT;
            sl->ca[s] = dynrm / dfnrm;
        }   // else: x == v stays as it is
    } else if (ph == PH_DT0) {
        const float est = absh * sqrtf(r0 * (1.f / NZ));
        if (0.1f * absh < hmax * sqrtf(est)) absh = fmaxf(0.1f * absh / sqrtf(est), hmin);
        else absh = hmax;
        sl->first[s] = 0;
        to_begin = true;
    } else {   // PH_STAGE
        const int m = sl->m[s], j = sl->j[s];
        if (j < m) {
            const float* c = P.tab + ((size_t)m * (MMAX + 1) + (j + 1)) * TAB_STRIDE;
            const float h = sl->h[s];
            sl->j[s] = j + 1;
            sl->ca[s] = c[0]; sl->cb[s] = c[1]; sl->cc[s] = c[2]; sl->cd[s] = h * c[3]; sl->ce[s] = h * c[4];
            act |= ACT_STAGE;
        } else {
            const float err = sqrtf(r0 * (1.f / NZ));
            const float h = sl->h[s];
            sl->vnrm2[s] = r2;
            if (!isfinite(err)) { sl->status[s] = ST_NAN; retire = true; }
            else if (err > 1.0f) {
                sl->nrej[s] += 1;
                absh = 0.8f * absh / cbrtf(err);
                if (absh < hmin) { sl->status[s] = ST_DT_UNDERFLOW; retire = true; }
                else if (sl->nacc[s] + sl->nrej[s] >= P.S.max_iters) { sl->status[s] = ST_MAXITERS; retire = true; }
                else {
                    const int newspc = !sl->jacatt[s];
                    sl->newspc[s] = newspc;
                    if (newspc) to_rho = true; else to_begin = true;
                }
            } else if (!(xmax <= bound)) {
                sl->status[s] = ST_NAN; retire = true;   // an accepted state outside the operand range of the split: report, never guess
            } else {
                const int nacc = sl->nacc[s] + 1;
                sl->nacc[s] = nacc;
                act |= AF_ACCEPT;
                sl->o_t0[s] = tcur; sl->o_h[s] = h; sl->o_row[s] = nacc - 1;
                tcur = sl->last[s] ? tf : tcur + h;
                sl->o_tn[s] = tcur; sl->t[s] = tcur;
                {
                    int si = sl->si[s];
                    sl->o_si[s] = si;
                    while (si < P.S.n_save && P.S.saveat[si] <= tcur) ++si;
                    sl->si[s] = si;
                }
                sl->ynrm2[s] = r1;
                sl->jacatt[s] = 0;
                const int nstsig = (sl->nstsig[s] + 1) % 25;
                sl->nstsig[s] = nstsig;
                const int newspc = nstsig == 0;
                sl->newspc[s] = newspc;
                float fac = 10.f;
                if (nacc == 1) {
                    const float t2 = cbrtf(err);
                    if (0.8f < fac * t2) fac = 0.8f / t2;
                } else {
                    const float t1 = 0.8f * absh * cbrtf(sl->errold[s]);
                    const float e3 = cbrtf(err);
                    const float t2 = sl->hlast[s] * e3 * e3;
                    if (t1 < fac * t2) fac = t1 / t2;
                }
                absh = fmaxf(0.1f, fac) * absh;
                absh = fmaxf(hmin, fminf(hmax, absh));
                sl->errold[s] = err; sl->hlast[s] = h;
                if (tcur >= tf) retire = true;
                else if (nacc + sl->nrej[s] >= P.S.max_iters) { sl->status[s] = ST_MAXITERS; retire = true; }
                else if (newspc) to_rho = true;
                else to_begin = true;
            }
        }
    }
    if (to_rho) {   // start a power iteration at (y_n, f_n): v <- ca*y_n + cb*v + cc
        const float ynrm = sqrtf(sl->ynrm2[s]), vnrm = sqrtf(sl->vnrm2[s]);
        const float sqrtu = 3.4526698e-04f;   // sqrt(eps_f32)
        float ca, cb, cc = 0.f, dynrm;
        if (ynrm != 0.f && vnrm != 0.f) { dynrm = ynrm * sqrtu; ca = 1.f; cb = dynrm / vnrm; }
        else if (ynrm != 0.f) { dynrm = ynrm * sqrtu; ca = 1.f + sqrtu; cb = 0.f; }
        else if (vnrm != 0.f) { dynrm = uround; ca = 0.f; cb = dynrm / vnrm; }
        else { dynrm = uround; ca = 0.f; cb = 0.f; cc = dynrm * 0.17677669529663687f; }   // 1/sqrt(32)
        sl->dynrm[s] = dynrm; sl->ca[s] = ca; sl->cb[s] = cb; sl->cc[s] = cc;
        sl->it[s] = 0; sl->sigma[s] = 0.f; sl->env[s] = 0.f; sl->smax[s] = 0.f;
        sl->phase[s] = PH_RHO;
        act |= ACT_RHO_INIT;
    }
    if (to_after_rho) {
        if (sl->first[s]) {
            absh = hmax;
            const float sprad = sl->sprad[s];
            if (sprad * absh > 1.f) absh = 1.f / sprad;
            absh = fmaxf(absh, hmin);
            sl->ca[s] = absh;
            sl->phase[s] = PH_DT0;
            act |= ACT_DT0;
        } else to_begin = true;
    }
    if (to_begin) {
        const float sprad = sl->sprad[s];
        int last = 0;
        if (1.1f * absh >= tf - tcur) { absh = tf - tcur; last = 1; }
        int m = 1 + (int)sqrtf(1.54f * absh * sprad + 1.f);
        if (m > P.mmax) { m = P.mmax; absh = (float)(m * m - 1) / (1.54f * sprad); last = 0; }
        sl->last[s] = last; sl->m[s] = m; sl->j[s] = 1; sl->h[s] = absh;
        sl->ca[s] = absh * P.tab[((size_t)m * (MMAX + 1) + 1) * TAB_STRIDE + 3];
        sl->phase[s] = PH_STAGE;
        act |= ACT_BEGIN;
    }
    if (retire) act = (act & ~ACT_MASK) | ACT_REFILL;
    sl->absh[s] = absh;
    sl->act[s] = act;
}

// RECORD: forward pass of nfc_loss_grad with the stabilised integrator -- residuals at the save points, their squared sum, and the
// adjoint tape: per accepted step u_n and the cubic-Hermite dense output in the power basis of the Tsit5 tape (c1 = f_n, c2, c3, c4 = 0),
// so the backward kernels (continuous adjoint: integrator-agnostic by construction) read it unchanged.
template <class N, bool WS, int NW, bool RECORD = false>
__global__ void __launch_bounds__(NW * 32, 1) solve_rkc_kernel(const Params P) {
    extern __shared__ __align__(16) float smem[];
    __shared__ uint64_t bar;
    __shared__ Slots slots[NW];
    const float* sW = WS ? smem : P.S.wpack;
    if (WS) stage_weights(smem, P.S.wpack, N::P_TOTAL, &bar);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float* ws = smem + (WS ? N::P_TOTAL : 0) + warp * N::S_TOTAL;
    Slots* sl = &slots[warp];
    const float reltol = P.S.reltol, abstol = P.S.abstol;

    float yn[CPW], fn[CPW], v[CPW], yjm2[CPW], x[CPW], f[CPW];
    double lsum = 0.0;
#pragma unroll
    for (int c = 0; c < CPW; ++c) {
        fn[c] = 0.f; v[c] = 0.f; yjm2[c] = 0.f;
        fill<RECORD>(P, sl, c, yn[c], x[c], lane, &lsum);
    }
    __syncwarp();

    while (true) {
        bool any = false;
#pragma unroll
        for (int c = 0; c < CPW; ++c) any |= sl->col[c] >= 0;
        if (!any) break;
        rhs_tile<N>(sW, ws, x, f, sl->bot, sl->top, sl->kca, P.S.pref, lane);

        // ---- (B) reductions; lane c keeps slot c's ----
        float r0 = 0.f, r1 = 0.f, r2 = 0.f;
#pragma unroll
        for (int c = 0; c < CPW; ++c) {
            if (sl->col[c] < 0) continue;
            const int ph = sl->phase[c];
            float a0 = 0.f, a1 = 0.f, a2 = 0.f;
            if (ph == PH_F0) {
                a1 = warp_sum(x[c] * x[c]); a2 = warp_sum(f[c] * f[c]);
            } else if (ph == PH_RHO) {
                const float d = f[c] - fn[c];
                a0 = warp_sum(d * d);
            } else if (ph == PH_DT0) {
                const float d = (f[c] - fn[c]) / (abstol + reltol * fabsf(yn[c]));
                a0 = warp_sum(d * d);
            } else if (sl->j[c] >= sl->m[c]) {
                const float est = 0.8f * (yn[c] - x[c]) + (0.4f * sl->h[c]) * (fn[c] + f[c]);
                const float d = est / (abstol + reltol * fmaxf(fabsf(x[c]), fabsf(yn[c])));
                a0 = warp_sum(d * d); a1 = warp_sum(x[c] * x[c]); a2 = warp_sum(v[c] * v[c]);
            }
            if (lane == c) { r0 = a0; r1 = a1; r2 = a2; }
        }
        __syncwarp();
        // ---- (C) scalar transitions, slot = lane ----
        if (lane < CPW && sl->col[lane] >= 0) transition(P, sl, lane, r0, r1, r2, 0.f, __int_as_float(0x7f800000));
        __syncwarp();
        // ---- (D) vector updates ----
#pragma unroll
        for (int c = 0; c < CPW; ++c) {
            if (sl->col[c] < 0) continue;
            const int act = sl->act[c];
            if (act & AF_SETF0) { fn[c] = f[c]; v[c] = f[c]; }
            if (act & AF_ACCEPT) {
                // cubic Hermite on (y_n, f_n, y_{n+1}, f_{n+1}) at every save time in (t, t+h]
                const long long col = sl->col[c];
                const float t0 = sl->o_t0[c], h = sl->o_h[c], tn = sl->o_tn[c];
                const float d = (x[c] - yn[c]) / h;
                const float c2 = 3.f * d - 2.f * fn[c] - f[c], c3 = -2.f * d + fn[c] + f[c];
                if (RECORD) {
                    const int n = sl->o_row[c];
                    if (n < tape_capacity(P.S, col)) {
                        float* e = P.S.tape + (tape_row0(P.S, col) + n) * TAPE_STRIDE;
                        e[lane] = yn[c]; e[NZ + lane] = fn[c]; e[2 * NZ + lane] = c2; e[3 * NZ + lane] = c3; e[4 * NZ + lane] = 0.f;
                        if (lane == 0) { e[TAPE_T] = t0; e[TAPE_H] = h; }
                    }
                }
                float sse = 0.f;
                for (int si = sl->o_si[c]; si < P.S.n_save; ++si) {
                    const float ts = P.S.saveat[si];
                    if (ts > tn) break;
                    const float th = fminf((ts - t0) / h, 1.0f);
                    const long long o = (col * P.S.n_save + si) * NZ + lane;
                    const float val = fmaf(h * th, fmaf(th, fmaf(th, c3, c2), fn[c]), yn[c]);
                    if (RECORD) { const float r = val - P.S.Ttrue[o]; P.S.resid[o] = r; sse = fmaf(r, r, sse); }
                    else if (P.S.Tout) P.S.Tout[o] = val;
                }
                if (RECORD) lsum += (double)warp_sum(sse);
                yn[c] = x[c]; fn[c] = f[c];
            }
            if (act & AF_RHO_DONE) v[c] -= yn[c];
            const int a = act & ACT_MASK;
            if (a == ACT_RHO_INIT) { v[c] = fmaf(sl->ca[c], yn[c], fmaf(sl->cb[c], v[c], sl->cc[c])); x[c] = v[c]; }
            else if (a == ACT_RHO_CONT) { v[c] = fmaf(f[c] - fn[c], sl->ca[c], yn[c]); x[c] = v[c]; }
            else if (a == ACT_DT0) { x[c] = fmaf(sl->ca[c], fn[c], yn[c]); }
            else if (a == ACT_BEGIN) { yjm2[c] = yn[c]; x[c] = fmaf(sl->ca[c], fn[c], yn[c]); }
            else if (a == ACT_STAGE) {
                const float yn1 = sl->ca[c] * x[c] + sl->cb[c] * yjm2[c] + sl->cc[c] * yn[c] + (sl->cd[c] * f[c] + sl->ce[c] * fn[c]);
                yjm2[c] = x[c]; x[c] = yn1;
            } else if (a == ACT_REFILL) {
                if (lane == 0) {
                    const int col = sl->col[c];
                    if (P.S.naccept) P.S.naccept[col] = sl->nacc[c];
                    if (P.S.nreject) P.S.nreject[col] = sl->nrej[c];
                    if (RECORD) {
                        if (sl->nacc[c] > tape_capacity(P.S, col) && sl->status[c] == ST_OK) sl->status[c] = ST_TAPE_OVERFLOW;
                        P.S.tape_len[col] = min(sl->nacc[c], tape_capacity(P.S, col));
                    }
                    if (P.S.status) P.S.status[col] = sl->status[c];
                    if (P.nfe) P.nfe[col] = sl->nfe[c];
                    atomicAdd(&P.S.counters[CTR_ACCEPT], (unsigned long long)sl->nacc[c]);
                    atomicAdd(&P.S.counters[CTR_REJECT], (unsigned long long)sl->nrej[c]);
                    atomicAdd(&P.S.counters[CTR_NFE], (unsigned long long)sl->nfe[c]);
                }
                __syncwarp();
                fn[c] = 0.f; v[c] = 0.f; yjm2[c] = 0.f;
                fill<RECORD>(P, sl, c, yn[c], x[c], lane, &lsum);
            }
        }
        __syncwarp();
    }
    if (RECORD && lane == 0) atomicAdd(P.S.loss_sum, lsum);
}

}  // namespace rkc
}  // namespace nfc
