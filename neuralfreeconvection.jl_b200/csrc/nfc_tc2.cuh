// nfc_tc2.cuh -- the tensor-core Tsit5 solve with TWO 128-column tiles in flight per CTA (FP16x2 split only).
//
// solve_tc_kernel (nfc_tc.cuh) owns one tile: its tensor pipe idles while the column threads run a layer's epilogue and the
// column threads idle while the MMAs of the next layer run (ncu: tensor pipe 27 % busy, issue slots 42 %).  Here the same 512
// column threads own two tiles, X and Y, and alternate between them, so that the MMAs of one tile run under the epilogue of
// the other:   P0(X) P0(Y) | E1(X) E1(Y) | E2(X) E2(Y) | E3(X) P0(X) | E3(Y) P0(Y) | ...      (P0 = stage input, E = epilogue)
// and the MMA warp issues  X.L1 Y.L1 X.L2 Y.L2 X.L3 Y.L3 ...  strictly alternating.
//
// What makes two tiles fit (one tile of nfc_tc.cuh uses 448 of the 512 tensor-memory columns):
//   * the A operand of the next layer is written IN PLACE over the thread's own slice of the accumulator: thread (column c,
//     quarter q) owns D columns [32q, 32q+32) = hidden units 32q..32q+31 and writes their FP16 hi parts to [32q, 32q+16) and
//     lo parts to [32q+16, 32q+32) (the MMA's A address is free per k-step), so no thread ever overwrites another thread's
//     accumulator and no barrier is needed between the loads and the stores.  The layer-1 operand (32 levels) goes to the
//     upper half [64, 128) of the output layer's region, which the N = 64 output accumulator does not touch;
//   * an MMA reads A from one 128-column region and writes D to another, so a tile needs two regions while its MMA runs and
//     one otherwise; with the strict X/Y alternation THREE regions serve both tiles (tile T, layer l: A = R[(l + 2T) % 3],
//     D = R[(l + 1 + 2T) % 3]; the region an MMA frees is the next MMA's accumulator -- the tensor pipe executes in issue order);
//   * of the ten stored stage vectors (k2..k6 of two tiles) four live in the remaining 128 tensor-memory columns (k2, k3) and
//     six in shared memory (k4..k6, 96 KB); the per-column scalars live in shared memory, owned by one thread per column (tile X: the
//     quarter-0 thread, tile Y: the quarter-1 thread, so the two tiles' controllers run side by side in different warps).
// Arithmetic is the one of solve_tc_kernel<SplitFP16x2, RECORD>, operation for operation (same MMA order, same combination order, same
// split and controller variants per mode), so the two kernels agree bit for bit -- tests/test_gpu_forward.py and
// tests/test_gpu_adjoint.py::test_record_pass_kernels_agree compare them.
#pragma once
#include "nfc_tc.cuh"

namespace nfc {
namespace tc2 {

using namespace tc;

constexpr int NT2 = NTHREADS;                   // 512 column threads + the MMA warpgroup (registers re-distributed with setmaxnreg as in nfc_tc.cuh:
                                                // registers are allocated per four warps, so a lone 17th warp would cost as much)
constexpr int TM_KT = 384;                      // 4 slots x 32 columns: k2, k3 of tile 0, k2, k3 of tile 1
constexpr int IMG_PAD = ((IMG16_BYTES + 127) / 128) * 128;
constexpr int KS_VEC_BYTES = 4 * 2 * TILE * 16; // one stage vector of one tile in shared memory: [quarter][half][column] float4
constexpr int KS_BYTES = 6 * KS_VEC_BYTES;      // k4, k5, k6 of both tiles
enum : int { CTL_ACCEPT = 1, CTL_NEWCOL = 2, CTL_RANGE = 4, CTL_LAST = 8 };

struct Cols {                                   // per-column scalars of one tile (written by the quarter-0 thread of the column)
    int col[TILE], si[TILE], nacc[TILE], nrej[TILE], status[TILE], ctl[TILE];
    float t[TILE], dt[TILE], qold[TILE], bot[TILE], top[TILE], kk[TILE], hs[TILE];
};
struct Shared2 {
    uint64_t bar_w, bar_a[2], bar_d[2];
    uint32_t tmem_base;
    int stop, qempty, ydead;                    // ydead: tile 1 has no live column and the queue is empty -> one-tile mode for the tail
    int anyT[2];
    float xch[2][2][TILE][4][2];                // [tile][buffer] first / last level of every quarter
    float part[2][TILE][4];                     // partial error norms
    Cols cs[2];
};
constexpr int SMEM2_BYTES = IMG_PAD + KS_BYTES + (int)sizeof(Shared2);
static_assert(SMEM2_BYTES <= 232448, "two-tile kernel: shared memory budget");

struct TileRegs { float u[8], k1[8], y[8], k7[8]; };

// clock64 phase timers exist only in the DBG instantiation (NFC_TC_DEBUG=1)
template <bool DBG> __device__ __forceinline__ long long tick() { return DBG ? clock64() : 0LL; }

// accumulator / operand regions: tile T, layer l reads A from region (l + 2T) % 3 and writes D to region (l + 1 + 2T) % 3
__host__ __device__ constexpr uint32_t region(int i) { return 128u * (uint32_t)(i % 3); }

// all MMAs of one Dense layer of one tile, in the order of tc::issue_layer16<LAYER, 0>, <LAYER, 1>, ... (sub-batches of two k-steps):
// the accumulation order decides the rounding, and the two kernels agree bit for bit
template <int LAYER>
__device__ __forceinline__ void issue2(uint32_t a_reg, uint32_t d_reg, uint32_t smem_img) {
    constexpr int K = LAYER == 0 ? 32 : 128, Nn = LAYER == 2 ? 64 : 128;
    constexpr uint32_t woff = LAYER == 0 ? OFF16_W1 : (LAYER == 1 ? OFF16_W2 : OFF16_W3);
    constexpr uint32_t wbytes = LAYER == 0 ? W1H_BYTES : W2H_BYTES;
    constexpr uint32_t idesc = make_idesc16(128, Nn);
    constexpr uint32_t kstep = (2u * Nn * 16u) >> 4;
    const uint64_t desc0 = make_bdesc(smem_img + woff, Nn);
    const uint32_t lo0 = (uint32_t)desc0, hi0 = (uint32_t)(desc0 >> 32);
#pragma unroll
    for (int sub = 0; sub < K / 32; ++sub) {
        if (LAYER == 2) {
#pragma unroll
            for (int sA = 1; sA >= 0; --sA)
#pragma unroll
                for (int j = 2 * sub; j < 2 * sub + 2; ++j) {
                    const uint32_t a = a_reg + 32 * (j >> 1) + 16 * sA + 8 * (j & 1);
                    if (sA == 1 && j == 0) mma_ts2<false>(d_reg, a, lo0 + kstep * j, hi0, idesc);
                    else mma_ts2<true>(d_reg, a, lo0 + kstep * j, hi0, idesc);
                }
        } else {
            constexpr int sa[3] = {1, 0, 0};      // (a split, b split): lo*hi, hi*lo, hi*hi
            constexpr int sb[3] = {0, 1, 0};
#pragma unroll
            for (int p = 0; p < 3; ++p)
#pragma unroll
                for (int j = 2 * sub; j < 2 * sub + 2; ++j) {
                    const uint32_t a = LAYER == 0 ? a_reg + 64 + 32 * sa[p] + 8 * j : a_reg + 32 * (j >> 1) + 16 * sa[p] + 8 * (j & 1);
                    const uint32_t blo = lo0 + ((sb[p] * wbytes) >> 4) + kstep * j;
                    if (p == 0 && j == 0) mma_ts2<false>(d_reg, a, blo, hi0, idesc);
                    else mma_ts2<true>(d_reg, a, blo, hi0, idesc);
                }
        }
    }
}

template <int T>
__device__ __forceinline__ void issue_tile_layer(int layer, uint32_t tb, uint32_t smem_img) {
    if (layer == 0) issue2<0>(tb + region(0 + 2 * T), tb + region(1 + 2 * T), smem_img);
    else if (layer == 1) issue2<1>(tb + region(1 + 2 * T), tb + region(2 + 2 * T), smem_img);
    else issue2<2>(tb + region(2 + 2 * T), tb + region(3 + 2 * T), smem_img);
}

__device__ __forceinline__ void mma_warp_loop2(Shared2* S, uint32_t smem_img, int lane) {
    if (lane == 0) {
        int pa0 = 0, pa1 = 0, layer = 0;
        while (true) {
            mbar_wait(&S->bar_a[0], pa0); pa0 ^= 1;
            if (*reinterpret_cast<volatile int*>(&S->stop)) break;
            fence_after_sync();
            // The operand addresses are re-derived from a base that is re-read in every iteration: as loop invariants the compiler
            // hoisted all 184 of them out of the loop and, with the 32 registers this warp keeps, spilled them -- one local-memory
            // load in front of every MMA made the issue slower than the tensor pipe.
            uint32_t tb = *reinterpret_cast<volatile uint32_t*>(&S->tmem_base), img = smem_img;
            asm volatile("" : "+r"(img));
            issue_tile_layer<0>(layer, tb, img);
            mma_commit(&S->bar_d[0]);
            if (!*reinterpret_cast<volatile int*>(&S->ydead)) {       // set before the arrival that starts a step, never cleared
                mbar_wait(&S->bar_a[1], pa1); pa1 ^= 1;
                fence_after_sync();
                tb = *reinterpret_cast<volatile uint32_t*>(&S->tmem_base);
                asm volatile("" : "+r"(img));
                issue_tile_layer<1>(layer, tb, img);
                mma_commit(&S->bar_d[1]);
            }
            layer = layer == 2 ? 0 : layer + 1;
        }
    }
    __syncwarp();
}

// ---- stage-vector storage: k2, k3 in tensor memory, k4..k6 in shared memory -----------------------------------------------
template <int T> __device__ __forceinline__ uint32_t kt_addr(uint32_t tl, int j, int q) { return tl + TM_KT + 32 * (2 * T + (j - 2)) + 8 * q; }
template <int T> __device__ __forceinline__ float4* ks_ptr(unsigned char* ks, int j, int q, int c) {
    return reinterpret_cast<float4*>(ks + (T * 3 + (j - 4)) * KS_VEC_BYTES) + (q * 2) * TILE + c;
}
template <int T> __device__ __forceinline__ void ks_load(unsigned char* ks, int j, int q, int c, float (&k)[8]) {
    const float4* p = ks_ptr<T>(ks, j, q, c);
    const float4 a = p[0], b = p[TILE];
    k[0] = a.x; k[1] = a.y; k[2] = a.z; k[3] = a.w; k[4] = b.x; k[5] = b.y; k[6] = b.z; k[7] = b.w;
}

// acc = sum_{j = 1 .. nk} coef[j-1] k_j in the operation order of tc::combine_n
template <int T>
__device__ __forceinline__ void combine(uint32_t tl, unsigned char* ks, int q, int c, const float (&k1)[8], const float* coef, int nk, float (&acc)[8]) {
    float kb[8], kc[8];
    scal8(acc, coef[0], k1);
    if (nk >= 2) {                                   // nk is uniform over the CTA: tcgen05.ld is warp-collective
        tmem_ld8f(kt_addr<T>(tl, 2, q), kb);
        if (nk >= 3) tmem_ld8f(kt_addr<T>(tl, 3, q), kc);
        wait_ld();
        axpy8(acc, coef[1], kb);
        if (nk >= 3) axpy8(acc, coef[2], kc);
    }
    if (nk >= 4) { ks_load<T>(ks, 4, q, c, kb); axpy8(acc, coef[3], kb); }
    if (nk >= 5) { ks_load<T>(ks, 5, q, c, kb); axpy8(acc, coef[4], kb); }
    if (nk >= 6) { ks_load<T>(ks, 6, q, c, kb); axpy8(acc, coef[5], kb); }
}

// ---- P0: stage input y = u + h sum a_sj k_j, published as the layer-1 A operand ------------------------------------------------
template <int T>
__device__ __forceinline__ void stage_input(Shared2* S, unsigned char* ks, const float* sb, TileRegs& R, int s, uint32_t tl, int q, int c) {
    float acc[8];
    combine<T>(tl, ks, q, c, R.k1, c_tsit5_a[s], s + 1, acc);
    const float hs = S->cs[T].hs[c];
#pragma unroll
    for (int i = 0; i < 8; i += 2) fma2s(R.y[i], R.y[i + 1], acc[i], acc[i + 1], hs, R.u[i], R.u[i + 1]);
    S->xch[T][s & 1][c][q][0] = R.y[0];
    S->xch[T][s & 1][c][q][1] = R.y[7];
    const float s_a0 = sb[291];
    uint32_t ph[4], pl[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float v0 = fminf(fmaxf(R.y[2 * i], -SplitFP16x2::INPUT_BOUND), SplitFP16x2::INPUT_BOUND);
        const float v1 = fminf(fmaxf(R.y[2 * i + 1], -SplitFP16x2::INPUT_BOUND), SplitFP16x2::INPUT_BOUND);
        float s0, s1;
        mul2s(s0, s1, v0, v1, s_a0);
        split2h(s0, s1, ph[i], pl[i]);
    }
    const uint32_t a = tl + region(2 * T) + 64 + 4 * q;
    tmem_st4(a, ph);
    tmem_st4(a + 32, pl);
    wait_st();
    fence_before_sync();
    warp_arrive(&S->bar_a[T]);
}

// ---- E1 / E2: hidden-layer epilogue, in place ---------------------------------------------------------------------------------
template <int T, bool DBG, bool FOLD>
__device__ __forceinline__ void hidden_epilogue(Shared2* S, const float* sb, int l, int& pd, uint32_t tl, int q, long long& tw) {
    { const long long a = tick<DBG>(); mbar_wait(&S->bar_d[T], pd); pd ^= 1; tw += tick<DBG>() - a; }
    fence_after_sync();
    const uint32_t d = tl + (T == 0 ? (l == 0 ? region(1) : region(2)) : (l == 0 ? region(0) : region(1))) + 32 * q;
    const float* b = sb + 128 * l + 32 * q;
    const float cl = sb[288 + l];
    uint32_t r[32];
    tmem_ld16(d, *reinterpret_cast<uint32_t(*)[16]>(&r[0]));
    tmem_ld16(d + 16, *reinterpret_cast<uint32_t(*)[16]>(&r[16]));
    wait_ld();
#pragma unroll
    for (int ch = 0; ch < 2; ++ch) {
        uint32_t ph[8], pl[8];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float4 bb = *reinterpret_cast<const float4*>(b + 16 * ch + 4 * i);
            float v0, v1, v2, v3;
            fma2s(v0, v1, __uint_as_float(r[16 * ch + 4 * i]), __uint_as_float(r[16 * ch + 4 * i + 1]), cl, bb.x, bb.y);
            fma2s(v2, v3, __uint_as_float(r[16 * ch + 4 * i + 2]), __uint_as_float(r[16 * ch + 4 * i + 3]), cl, bb.z, bb.w);
            if (FOLD) {                                              // relu folded into the conversions (tc::split2h_relu); not in the RECORD pass
                split2h_relu(v0, v1, ph[2 * i], pl[2 * i]);
                split2h_relu(v2, v3, ph[2 * i + 1], pl[2 * i + 1]);
            } else {
                split2h(fmaxf(v0, 0.f), fmaxf(v1, 0.f), ph[2 * i], pl[2 * i]);
                split2h(fmaxf(v2, 0.f), fmaxf(v3, 0.f), ph[2 * i + 1], pl[2 * i + 1]);
            }
        }
        tmem_st8(d + 8 * ch, ph);              // units 32q + 16ch .. : k-step 2q + ch of the next layer, hi
        tmem_st8(d + 16 + 8 * ch, pl);         //                                                        lo
    }
    wait_st();
    fence_before_sync();
    warp_arrive(&S->bar_a[T]);
}

// ---- E3: output layer -> f(y) -> k_{s+2} ------------------------------------------------------------------------------------------
template <int T, bool DBG>
__device__ __forceinline__ void output_epilogue(Shared2* S, unsigned char* ks, const float* sb, TileRegs& R, int s, int& pd, uint32_t tl, int q, int c, float pref, long long& tw) {
    const float ylo = q > 0 ? S->xch[T][s & 1][c][q - 1][1] : 0.f, yhi = q < 3 ? S->xch[T][s & 1][c][q + 1][0] : 0.f;
    const float bot = S->cs[T].bot[c], top = S->cs[T].top[c], kk = S->cs[T].kk[c];
    { const long long a = tick<DBG>(); mbar_wait(&S->bar_d[T], pd); pd ^= 1; tw += tick<DBG>() - a; }
    fence_after_sync();
    float nn[9];
    {
        const uint32_t d = tl + region(2 * T);
        uint32_t ra[8], rb[8], a0 = 0, b0 = 0;
        tmem_ld8(d + 8 * q, ra);
        tmem_ld8(d + 32 + 8 * q, rb);
        if (q) { tmem_ld1(d + 8 * q - 1, a0); tmem_ld1(d + 32 + 8 * q - 1, b0); }
        wait_ld();
        const float* b3 = sb + 256;
        const float inv3 = sb[290];
        nn[0] = q ? fmaf(__uint_as_float(b0) + __uint_as_float(a0), inv3, b3[8 * q - 1]) : 0.f;
#pragma unroll
        for (int i = 0; i < 8; i += 2) {
            float t0, t1;
            add2(t0, t1, __uint_as_float(rb[i]), __uint_as_float(rb[i + 1]), __uint_as_float(ra[i]), __uint_as_float(ra[i + 1]));
            fma2s(nn[i + 1], nn[i + 2], t0, t1, inv3, b3[8 * q + i], b3[8 * q + i + 1]);
        }
    }
    fence_before_sync();
    faces_to_f(R.y, ylo, yhi, nn, bot, top, kk, pref, q, R.k7);
    if (s <= 1) {                                    // uniform over the CTA
        tmem_st8f(kt_addr<T>(tl, s + 2, q), R.k7);
        wait_st();
    } else if (s <= 4) {
        float4* p = ks_ptr<T>(ks, s + 2, q, c);
        p[0] = make_float4(R.k7[0], R.k7[1], R.k7[2], R.k7[3]);
        p[TILE] = make_float4(R.k7[4], R.k7[5], R.k7[6], R.k7[7]);
    }
}

// ---- end of a step ------------------------------------------------------------------------------------------------------------
struct StepOld { float t, hs; int col, si, ctl, nacc; };

template <int T>
__device__ __forceinline__ void step_error(Shared2* S, unsigned char* ks, const SolveParams& P, TileRegs& R, StepOld& o, uint32_t tl, int q, int c) {
    o.t = S->cs[T].t[c]; o.hs = S->cs[T].hs[c]; o.col = S->cs[T].col[c]; o.si = S->cs[T].si[c]; o.ctl = S->cs[T].ctl[c]; o.nacc = S->cs[T].nacc[c];
    const float bt[6] = {NFC_BT1, NFC_BT2, NFC_BT3, NFC_BT4, NFC_BT5, NFC_BT6};
    float e[8], part = 0.f;
    combine<T>(tl, ks, q, c, R.k1, bt, 6, e);
#pragma unroll
    for (int i = 0; i < 8; ++i) part = err_term(part, e[i], R.k7[i], o.hs, R.u[i], R.y[i], P.abstol, P.reltol);
    S->part[T][c][q] = part;
    float m = fabsf(R.y[0]);
#pragma unroll
    for (int i = 1; i < 8; ++i) m = fmaxf(m, fabsf(R.y[i]));
    if (!(m <= SplitFP16x2::INPUT_BOUND)) atomicOr(&S->cs[T].ctl[c], (int)CTL_RANGE);
}

// quarter-0 thread of column c: controller, retire / refill, next step size (scalar state in shared memory)
template <int T, bool RECORD>
__device__ __forceinline__ void step_control(Shared2* S, const SolveParams& P, int c) {
    Cols& C = S->cs[T];
    int col = C.col[c];
    int newctl = 0;
    bool done = col < 0;
    if (col >= 0) {
        const int ctl = C.ctl[c];
        const float t = C.t[c], hs = C.hs[c];
        int status = C.status[c], nacc = C.nacc[c], nrej = C.nrej[c];
        const bool last = ctl & CTL_LAST;
        const float ee = fast_sqrt(((S->part[T][c][0] + S->part[T][c][1]) + (S->part[T][c][2] + S->part[T][c][3])) * (1.f / NZ));
        bool accept;
        float dtnew;
        if (!isfinite(ee)) { status = ST_NAN; accept = false; dtnew = hs; }
        else if (P.fixed_dt > 0.f) { accept = true; dtnew = P.fixed_dt; }
        else pi_controller(ee, C.qold[c], hs, accept, dtnew);      // as in tc::solve_tc_kernel (solve and RECORD pass: one arithmetic)
        if (accept && (ctl & CTL_RANGE)) { status = ST_RANGE; accept = false; }     // accepted state outside the operand range of the split
        if (accept) {
            const float tn = last ? P.tf : t + hs;
            int si = C.si[c];
            while (si < P.n_save && !(P.saveat[si] > tn)) ++si;
            C.si[c] = si;
            C.t[c] = tn;
            if (P.fixed_dt <= 0.f) C.qold[c] = fmaxf(ee, 1e-4f);
            if (RECORD && nacc >= tape_capacity(P, col)) status = ST_TAPE_OVERFLOW;      // no tape row left for this step
            ++nacc;
            if (tn >= P.tf) done = true;
            newctl |= CTL_ACCEPT;
        } else {
            ++nrej;
        }
        C.dt[c] = dtnew;
        if (status == ST_OK && !done) {
            if (dtnew < 1e-12f) status = ST_DT_UNDERFLOW;
            else if (nacc + nrej >= P.max_iters) status = ST_MAXITERS;
        }
        if (status != ST_OK) done = true;
        C.status[c] = status; C.nacc[c] = nacc; C.nrej[c] = nrej;
        if (done) {
            if (P.naccept) P.naccept[col] = nacc;
            if (P.nreject) P.nreject[col] = nrej;
            if (P.status) P.status[col] = status;
            if (P.work) P.work[col] = nacc + nrej;
            if (RECORD) P.tape_len[col] = min(nacc, tape_capacity(P, col));
            atomicAdd(&P.counters[CTR_ACCEPT], (unsigned long long)nacc);
            atomicAdd(&P.counters[CTR_REJECT], (unsigned long long)nrej);
        }
    }
    if (done) {
        int nc = -1;
        if (!*reinterpret_cast<volatile int*>(&S->qempty)) {
            const long long nx = (long long)atomicAdd(&P.counters[CTR_QUEUE], 1ULL);
            if (nx < P.B) nc = P.order ? P.order[nx] : (int)nx; else S->qempty = 1;
        }
        if (nc >= 0 || col >= 0) newctl |= CTL_NEWCOL;
        col = nc;
        C.col[c] = nc;
        C.t[c] = 0.f; C.qold[c] = 1e-4f; C.nacc[c] = 0; C.nrej[c] = 0; C.status[c] = ST_OK;
        if (nc >= 0) {
            C.dt[c] = P.dt_init[nc];
            C.bot[c] = P.colp[(long long)nc * 3 + 0]; C.top[c] = P.colp[(long long)nc * 3 + 1];
            C.kk[c] = P.use_kca ? P.colp[(long long)nc * 3 + 2] * (float)NZ : 0.f;
            int si = 0;
            while (si < P.n_save && P.saveat[si] <= 0.f) ++si;
            C.si[c] = si;
        } else {
            C.dt[c] = 0.f; C.bot[c] = 0.f; C.top[c] = 0.f; C.kk[c] = 0.f; C.si[c] = 0;
        }
    }
    float hs = 0.f;
    if (col >= 0) {
        const float t = C.t[c];
        hs = fminf(C.dt[c], P.tf - t);
        if (t + hs >= P.tf * (1.0f - 1e-6f)) { hs = P.tf - t; newctl |= CTL_LAST; }
        S->anyT[T] = 1;
    }
    C.hs[c] = hs;
    C.ctl[c] = newctl;
}

// all four threads of a column: dense output of the accepted step, state update, load of a new column
// (RECORD: residuals, tape row and squared-residual sum instead of the trajectory, as in tc::solve_tc_kernel<SP, true>)
template <bool RECORD>
__device__ __forceinline__ void save_point(const SolveParams& P, long long o2i, const float (&v)[8], double& lsum) {
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
        float* o2 = P.Tout + o2i;          // streaming stores (see tc::solve_tc_kernel)
        __stcs(reinterpret_cast<float4*>(o2), make_float4(v[0], v[1], v[2], v[3]));
        __stcs(reinterpret_cast<float4*>(o2 + 4), make_float4(v[4], v[5], v[6], v[7]));
    }
}

template <int T, bool RECORD>
__device__ __forceinline__ void step_finish(Shared2* S, unsigned char* ks, const SolveParams& P, TileRegs& R, const StepOld& o, uint32_t tl, int q, int c, double& lsum) {
    // dense output in power basis, u(t + th h) = u + h th (k1 + th (c2 + th (c3 + th c4))): the loads are warp-collective
    float c2[8], c3[8], c4[8];
    {
        const float r2[6] = {-2.763706197274826f, 0.13169999999999998f, 3.9302962368947516f, -12.411077166933676f, 37.50931341651104f, -27.896526289197286f};
        const float r3[6] = {2.9132554618219126f, -0.2234f, -5.941033872131505f, 30.33818863028232f, -88.1789048947664f, 65.09189467479366f};
        const float r4[6] = {-1.0530884977290216f, 0.1017f, 2.490627285651253f, -16.548102889244902f, 47.37952196281928f, -34.87065786149661f};
        // every stage vector is loaded once for the three combinations (each in the operation order of tc::dense_partial)
        float kb[8], kc[8];
        tmem_ld8f(kt_addr<T>(tl, 2, q), kb);
        tmem_ld8f(kt_addr<T>(tl, 3, q), kc);
        scal8(c2, r2[0], R.k1); scal8(c3, r3[0], R.k1); scal8(c4, r4[0], R.k1);
        wait_ld();
        axpy8(c2, r2[1], kb); axpy8(c3, r3[1], kb); axpy8(c4, r4[1], kb);
        axpy8(c2, r2[2], kc); axpy8(c3, r3[2], kc); axpy8(c4, r4[2], kc);
#pragma unroll
        for (int j = 4; j <= 6; ++j) {
            ks_load<T>(ks, j, q, c, kb);
            axpy8(c2, r2[j - 1], kb); axpy8(c3, r3[j - 1], kb); axpy8(c4, r4[j - 1], kb);
        }
        dense_k7(R.k7, c2, c3, c4);
    }
    const int ctl = S->cs[T].ctl[c];
    if (o.col >= 0 && (ctl & CTL_ACCEPT)) {
        const float tn = (o.ctl & CTL_LAST) ? P.tf : o.t + o.hs;
        int si = o.si;
        if (RECORD && o.nacc < tape_capacity(P, o.col)) {
            float* e = P.tape + (tape_row0(P, o.col) + o.nacc) * TAPE_STRIDE + 8 * q;
#pragma unroll
            for (int hlf = 0; hlf < 2; ++hlf) {
                const int i = 4 * hlf;
                *reinterpret_cast<float4*>(e + i) = make_float4(R.u[i], R.u[i + 1], R.u[i + 2], R.u[i + 3]);
                *reinterpret_cast<float4*>(e + NZ + i) = make_float4(R.k1[i], R.k1[i + 1], R.k1[i + 2], R.k1[i + 3]);
                *reinterpret_cast<float4*>(e + 2 * NZ + i) = make_float4(c2[i], c2[i + 1], c2[i + 2], c2[i + 3]);
                *reinterpret_cast<float4*>(e + 3 * NZ + i) = make_float4(c3[i], c3[i + 1], c3[i + 2], c3[i + 3]);
                *reinterpret_cast<float4*>(e + 4 * NZ + i) = make_float4(c4[i], c4[i + 1], c4[i + 2], c4[i + 3]);
            }
            if (q == 0) { e[TAPE_T] = o.t; e[TAPE_H] = o.hs; }
        }
        while (si < P.n_save) {
            const float ts = P.saveat[si];
            if (ts > tn) break;
            const float th = fminf(fast_div(ts - o.t, o.hs), 1.0f);
            float v[8];
            dense_eval(v, o.hs, th, R.u, R.k1, c2, c3, c4);
            save_point<RECORD>(P, ((long long)o.col * P.n_save + si) * NZ + 8 * q, v, lsum);
            ++si;
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) { R.u[i] = R.y[i]; R.k1[i] = R.k7[i]; }
    }
    if (ctl & CTL_NEWCOL) {
        const int nc = S->cs[T].col[c];
        if (nc >= 0) {
            const long long off = (long long)nc * NZ + 8 * q;
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const float4 a = *reinterpret_cast<const float4*>(P.T0 + off + 4 * i), b = *reinterpret_cast<const float4*>(P.k1_init + off + 4 * i);
                R.u[4 * i] = a.x; R.u[4 * i + 1] = a.y; R.u[4 * i + 2] = a.z; R.u[4 * i + 3] = a.w;
                R.k1[4 * i] = b.x; R.k1[4 * i + 1] = b.y; R.k1[4 * i + 2] = b.z; R.k1[4 * i + 3] = b.w;
            }
            int si = 0;
            while (si < P.n_save && P.saveat[si] <= 0.f) {
                save_point<RECORD>(P, ((long long)nc * P.n_save + si) * NZ + 8 * q, R.u, lsum);
                ++si;
            }
        } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) { R.u[i] = 0.f; R.k1[i] = 0.f; }
        }
    }
}

template <bool DBG, bool RECORD = false>
__global__ void __launch_bounds__(NT2, 1) solve_tc2_kernel(const unsigned char* __restrict__ gimg, const SolveParams P) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned char* ks = smem_raw + IMG_PAD;
    Shared2* S = reinterpret_cast<Shared2*>(smem_raw + IMG_PAD + KS_BYTES);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        mbar_init(&S->bar_w, 1);
        mbar_init(&S->bar_a[0], NCW); mbar_init(&S->bar_a[1], NCW);
        mbar_init(&S->bar_d[0], 1); mbar_init(&S->bar_d[1], 1);
        S->stop = 0; S->qempty = 0; S->ydead = 0; S->anyT[0] = 0; S->anyT[1] = 0;
    }
    __syncthreads();
    if (warp == NCT / 32) tmem_alloc(&S->tmem_base, TM_COLS);
    if (threadIdx.x == 0) {
        mbar_expect_tx(&S->bar_w, IMG16_BYTES);
        for (unsigned off = 0; off < (unsigned)IMG16_BYTES; off += 32768u) {
            const unsigned n = IMG16_BYTES - off < 32768u ? IMG16_BYTES - off : 32768u;
            bulk_g2s(smem_raw + off, gimg + off, n, &S->bar_w);
        }
    }
    if (threadIdx.x < NCT) {
        const int cc = threadIdx.x & (TILE - 1), tt = (threadIdx.x >> 7) & 1;
        if (threadIdx.x < 2 * TILE) {
            Cols& C = S->cs[tt];
            C.col[cc] = -1; C.si[cc] = 0; C.nacc[cc] = 0; C.nrej[cc] = 0; C.status[cc] = ST_OK; C.ctl[cc] = 0;
            C.t[cc] = 0.f; C.dt[cc] = 0.f; C.qold[cc] = 1e-4f; C.bot[cc] = 0.f; C.top[cc] = 0.f; C.kk[cc] = 0.f; C.hs[cc] = 0.f;
        }
    }
    mbar_wait(&S->bar_w, 0);
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t smem_img = smem_u32(smem_raw);
    const float* sb = reinterpret_cast<const float*>(smem_raw + OFF16_B);
    if (warp >= NCT / 32) {
        reg_dec<REGS_MMA>();
        if (warp == NCT / 32) mma_warp_loop2(S, smem_img, lane);
    } else {
        reg_inc<REGS_COL>();
        const int wq = warp & 3, q = warp >> 2, c = 32 * wq + lane;
        const uint32_t tl = S->tmem_base + ((uint32_t)(32 * wq) << 16);
        TileRegs X, Y;
#pragma unroll
        for (int i = 0; i < 8; ++i) { X.u[i] = X.k1[i] = X.y[i] = X.k7[i] = 0.f; Y.u[i] = Y.k1[i] = Y.y[i] = Y.k7[i] = 0.f; }
        int pdx = 0, pdy = 0;
        {   // the stage vectors in tensor memory must be finite before the first error estimate of an empty slot reads them
            const float zero8[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int j = 0; j < 4; ++j) tmem_st8f(tl + TM_KT + 32 * j + 8 * q, zero8);
            wait_st();
#pragma unroll
            for (int j = 4; j <= 6; ++j) {
                float4* p0 = ks_ptr<0>(ks, j, q, c); float4* p1 = ks_ptr<1>(ks, j, q, c);
                p0[0] = p0[TILE] = p1[0] = p1[TILE] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
        // phase timers of this thread (clock64): [0] step end, [1] stage inputs, [2] hidden epilogues, [3] output epilogues, [4..6] of which
        // waiting for the accumulator of layer 1 / 2 / 3, [7] stages
        long long tm[8] = {0, 0, 0, 0, 0, 0, 0, 0}, te[4] = {0, 0, 0, 0};   // te: step end = error estimate | barrier | control + barrier | finish
        bool ydead = false;
        double lsum = 0.0;                   // RECORD: this thread's sum of squared residuals
        while (true) {
            StepOld ox, oy;
            long long c0 = tick<DBG>();
            step_error<0>(S, ks, P, X, ox, tl, q, c);
            if (!ydead) step_error<1>(S, ks, P, Y, oy, tl, q, c);
            if (threadIdx.x == 0) { S->anyT[0] = 0; S->anyT[1] = 0; }
            long long e1 = tick<DBG>(); te[0] += e1 - c0;
            cta_bar();
            long long e2 = tick<DBG>(); te[1] += e2 - e1;
            if (q == 0) step_control<0, RECORD>(S, P, c);         // the two tiles' controllers run side by side in different warps
            else if (q == 1 && !ydead) step_control<1, RECORD>(S, P, c);
            cta_bar();
            e1 = tick<DBG>(); te[2] += e1 - e2;
            step_finish<0, RECORD>(S, ks, P, X, ox, tl, q, c, lsum);
            if (!ydead) step_finish<1, RECORD>(S, ks, P, Y, oy, tl, q, c, lsum);
            te[3] += tick<DBG>() - e1;
            const bool anyx = *reinterpret_cast<volatile int*>(&S->anyT[0]) != 0, anyy = !ydead && *reinterpret_cast<volatile int*>(&S->anyT[1]) != 0;
            if (!anyx && !anyy) break;
            if (!ydead && (!anyx || !anyy) && *reinterpret_cast<volatile int*>(&S->qempty)) {
                // One tile ran dry and nothing is left to refill it: finish the tail in one-tile mode (a stage of one tile costs
                // 8.5 K cycles, of two tiles 13 K).  If it is tile 0 that ran dry, tile 1 moves into its place first: at a step
                // boundary a tile's state is (u, k1) and the per-column scalars.
                if (!anyx) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) { X.u[i] = Y.u[i]; X.k1[i] = Y.k1[i]; }
                    if (q == 0) {
                        Cols& A = S->cs[0]; const Cols& B = S->cs[1];
                        A.col[c] = B.col[c]; A.si[c] = B.si[c]; A.nacc[c] = B.nacc[c]; A.nrej[c] = B.nrej[c]; A.status[c] = B.status[c]; A.ctl[c] = B.ctl[c];
                        A.t[c] = B.t[c]; A.dt[c] = B.dt[c]; A.qold[c] = B.qold[c]; A.bot[c] = B.bot[c]; A.top[c] = B.top[c]; A.kk[c] = B.kk[c]; A.hs[c] = B.hs[c];
                    }
                }
                ydead = true;
                if (threadIdx.x == 0) S->ydead = 1;
                cta_bar();
            }
            long long c1 = tick<DBG>(); tm[0] += c1 - c0; c0 = c1;
            // ---------------- one attempted Tsit5 step of both tiles ----------------
            stage_input<0>(S, ks, sb, X, 0, tl, q, c);
            if (!ydead) stage_input<1>(S, ks, sb, Y, 0, tl, q, c);
            c1 = tick<DBG>(); tm[1] += c1 - c0; c0 = c1;
#pragma unroll 1
            for (int s = 0; s < 6; ++s) {
#pragma unroll 1
                for (int l = 0; l < 2; ++l) {
                    hidden_epilogue<0, DBG, true>(S, sb, l, pdx, tl, q, tm[4 + l]);
                    if (!ydead) hidden_epilogue<1, DBG, true>(S, sb, l, pdy, tl, q, tm[4 + l]);
                }
                c1 = tick<DBG>(); tm[2] += c1 - c0; c0 = c1;
                output_epilogue<0, DBG>(S, ks, sb, X, s, pdx, tl, q, c, P.pref, tm[6]);
                c1 = tick<DBG>(); tm[3] += c1 - c0; c0 = c1;
                if (s < 5) stage_input<0>(S, ks, sb, X, s + 1, tl, q, c);
                c1 = tick<DBG>(); tm[1] += c1 - c0; c0 = c1;
                if (!ydead) {
                    output_epilogue<1, DBG>(S, ks, sb, Y, s, pdy, tl, q, c, P.pref, tm[6]);
                    c1 = tick<DBG>(); tm[3] += c1 - c0; c0 = c1;
                    if (s < 5) stage_input<1>(S, ks, sb, Y, s + 1, tl, q, c);
                    c1 = tick<DBG>(); tm[1] += c1 - c0; c0 = c1;
                }
                tm[7] += 1;
            }
        }
        if (RECORD) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) lsum += __shfl_xor_sync(FULL, lsum, o);
            if (lane == 0) atomicAdd(P.loss_sum, lsum);
        }
        if (DBG && !RECORD && P.resid && blockIdx.x == 0 && threadIdx.x == 0) {
            long long* dbg = reinterpret_cast<long long*>(P.resid);
#pragma unroll
            for (int i = 0; i < 8; ++i) dbg[i] = tm[i];
#pragma unroll
            for (int i = 0; i < 4; ++i) dbg[8 + i] = te[i];
        }
        // release the MMA warp (it waits for tile 0's next operand)
        cta_bar();
        if (threadIdx.x == 0) S->stop = 1;
        cta_bar();
        warp_arrive(&S->bar_a[0]);
    }
    fence_before_sync();
    __syncthreads();
    if (warp == NCT / 32) tmem_dealloc(S->tmem_base, TM_COLS);
}

}  // namespace tc2
}  // namespace nfc
