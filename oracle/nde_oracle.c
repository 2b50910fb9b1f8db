/* CPU ORACLE -- TEST / BASELINE INFRASTRUCTURE ONLY (never linked into the product library).
 *
 * Float32 C twin of oracle/nde_oracle.py: one column at a time with per-column gemv, the same
 * structure as the reference (one ODE solve per simulation, src/training.jl:46, BLAS on one
 * thread, train_free_convection_nde.jl:23), minus the reference's per-RHS-call allocations
 * (src/free_convection_nde.jl:30-36) -- i.e. a faster-than-reference CPU baseline ("port").
 * OpenMP over columns is optional (nthreads argument).
 *
 * PARITY UNPINNED: restates the Julia sources cited below; the reference ships no golden vectors.
 *
 *   RHS            src/free_convection_nde.jl:29-38, src/convective_adjustment_nde.jl:33-48
 *   theta layout   Flux.destructure (src/solve.jl:2): per Dense: weight out x in column-major, bias
 *   solve          src/solve.jl:4  reltol=1e-4, Tsit5 + PI controller + Hairer initial dt [UPSTREAM OrdinaryDiffEq 5.60.1]
 *   loss           src/training.jl:45-48  mean((yhat-y)^2)
 *   gradient       src/training.jl:70 (Zygote + DiffEqSensitivity) restated as the discrete adjoint of the
 *                  accepted Tsit5 step sequence incl. dense output
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define MAXL 8
#define MAXW 512

typedef struct {
    int Nz, conv_width, n_layers;
    int n_in[MAXL], n_out[MAXL], act[MAXL];
    int64_t w_off[MAXL], b_off[MAXL];
    int64_t conv_off, n_params;
} net_t;

static const float A21 = 0.161f;
static const float A31 = -0.008480655492356989f, A32 = 0.335480655492357f;
static const float A41 = 2.8971530571054935f, A42 = -6.359448489975075f, A43 = 4.3622954328695815f;
static const float A51 = 5.325864828439257f, A52 = -11.748883564062828f, A53 = 7.4955393428898365f, A54 = -0.09249506636175525f;
static const float A61 = 5.86145544294642f, A62 = -12.92096931784711f, A63 = 8.159367898576159f, A64 = -0.071584973281401f, A65 = -0.028269050394068383f;
static const float B1 = 0.09646076681806523f, B2 = 0.01f, B3 = 0.4798896504144996f, B4 = 1.379008574103742f, B5 = -3.290069515436081f, B6 = 2.324710524099774f;
static const float BT[7] = {-0.00178001105222577714f, -0.0008164344596567469f, 0.007880878010261995f, -0.1447110071732629f,
                            0.5823571654525552f, -0.45808210592918697f, 0.015151515151515152f};
static const float RI[7][4] = {
    {1.0f, -2.763706197274826f, 2.9132554618219126f, -1.0530884977290216f},
    {0.0f, 0.13169999999999998f, -0.2234f, 0.1017f},
    {0.0f, 3.9302962368947516f, -5.941033872131505f, 2.490627285651253f},
    {0.0f, -12.411077166933676f, 30.33818863028232f, -16.548102889244902f},
    {0.0f, 37.50931341651104f, -88.1789048947664f, 47.37952196281928f},
    {0.0f, -27.896526289197286f, 65.09189467479366f, -34.87065786149661f},
    {0.0f, 1.5f, -4.0f, 2.5f}};

static void a_row(int s, float *a) { /* a[j] for stage s (1-based 2..7), j = 0..s-2 */
    switch (s) {
    case 2: a[0] = A21; break;
    case 3: a[0] = A31; a[1] = A32; break;
    case 4: a[0] = A41; a[1] = A42; a[2] = A43; break;
    case 5: a[0] = A51; a[1] = A52; a[2] = A53; a[3] = A54; break;
    case 6: a[0] = A61; a[1] = A62; a[2] = A63; a[3] = A64; a[4] = A65; break;
    default: a[0] = B1; a[1] = B2; a[2] = B3; a[3] = B4; a[4] = B5; a[5] = B6; break;
    }
}

static int net_init(net_t *n, int Nz, int conv_width, int n_layers, const int *dims, const int *acts) {
    if (n_layers < 1 || n_layers > MAXL) return -1;
    if (Nz < 2 || Nz > 64) return -1;
    n->Nz = Nz; n->conv_width = conv_width; n->n_layers = n_layers;
    int64_t off = 0;
    n->conv_off = 0;
    if (conv_width) off += conv_width + 1;
    for (int l = 0; l < n_layers; ++l) {
        n->n_in[l] = dims[l]; n->n_out[l] = dims[l + 1]; n->act[l] = acts[l];
        if (dims[l] > MAXW || dims[l + 1] > MAXW) return -1;
        n->w_off[l] = off; off += (int64_t)dims[l] * dims[l + 1];
        n->b_off[l] = off; off += dims[l + 1];
    }
    n->n_params = off;
    if (dims[0] != Nz - (conv_width ? conv_width - 1 : 0) || dims[n_layers] != Nz - 1) return -1;
    return 0;
}

/* activations acts[l] = input of Dense l (l = 0..n_layers), acts[n_layers] = NN output; xin = raw T (conv input) */
typedef struct { float xin[MAXW]; float pre0[MAXW]; float a[MAXL + 1][MAXW]; } actbuf_t;

static void nn_forward(const net_t *n, const float *th, const float *T, actbuf_t *ab) {
    int n0 = n->n_in[0];
    if (n->conv_width) {
        int k = n->conv_width;
        const float *w = th + n->conv_off; float b = w[k];
        for (int i = 0; i < n->Nz; ++i) ab->xin[i] = T[i];
        for (int i = 0; i < n0; ++i) {
            float acc = 0.f;
            for (int j = 0; j < k; ++j) acc += w[j] * T[i + (k - 1 - j)];     /* flipped kernel (Flux Conv) */
            acc += b;
            ab->pre0[i] = acc;
            ab->a[0][i] = acc > 0.f ? acc : 0.f;
        }
    } else {
        for (int i = 0; i < n0; ++i) ab->a[0][i] = T[i];
    }
    for (int l = 0; l < n->n_layers; ++l) {
        int ni = n->n_in[l], no = n->n_out[l];
        const float *W = th + n->w_off[l], *b = th + n->b_off[l];
        float *y = ab->a[l + 1]; const float *x = ab->a[l];
        for (int o = 0; o < no; ++o) y[o] = b[o];
        for (int i = 0; i < ni; ++i) { float xi = x[i]; const float *Wc = W + (int64_t)i * no;
            for (int o = 0; o < no; ++o) y[o] += Wc[o] * xi; }
        if (n->act[l]) for (int o = 0; o < no; ++o) y[o] = y[o] > 0.f ? y[o] : 0.f;
    }
}

/* f(T): colp = (bottom, top, K_CA); pref = sigma_wT/sigma_T*tau/H */
static void rhs_eval(const net_t *n, const float *th, const float *T, const float *colp, float pref, float *out, actbuf_t *ab) {
    int N = n->Nz; float invdz = (float)N;
    nn_forward(n, th, T, ab);
    const float *nn = ab->a[n->n_layers];
    float Fk = colp[0];
    for (int k = 0; k < N; ++k) {
        float Fk1;
        if (k == N - 1) Fk1 = colp[1];
        else { float d = colp[2] * ((T[k + 1] - T[k]) * invdz); Fk1 = nn[k] - (d < 0.f ? d : 0.f); }
        out[k] = -pref * ((Fk1 - Fk) * invdz);
        Fk = Fk1;
    }
}

/* VJP: given kbar = dL/df, accumulates dL/dT into ubar (+=) and dL/dtheta into g (+=, double). */
static void rhs_vjp(const net_t *n, const float *th, const float *T, const float *colp, float pref, const float *kbar,
                    float *ubar, double *g, actbuf_t *ab) {
    int N = n->Nz; float invdz = (float)N; float c = pref * invdz;
    nn_forward(n, th, T, ab);
    float delta[MAXW], dprev[MAXW], tbar[MAXW];
    for (int k = 0; k < N; ++k) tbar[k] = 0.f;
    /* Fbar_k = c (kbar_k - kbar_{k-1}) for faces k = 1..N-1 ; NN output index k-1 */
    for (int k = 1; k < N; ++k) {
        float Fb = c * (kbar[k] - kbar[k - 1]);
        delta[k - 1] = Fb;
        float d = colp[2] * ((T[k] - T[k - 1]) * invdz);
        if (d < 0.f) { float s = colp[2] * invdz * Fb; tbar[k] -= s; tbar[k - 1] += s; }
    }
    for (int l = n->n_layers - 1; l >= 0; --l) {
        int ni = n->n_in[l], no = n->n_out[l];
        const float *W = th + n->w_off[l]; const float *x = ab->a[l]; const float *y = ab->a[l + 1];
        if (n->act[l]) for (int o = 0; o < no; ++o) if (!(y[o] > 0.f)) delta[o] = 0.f;
        double *gW = g + n->w_off[l], *gb = g + n->b_off[l];
        for (int o = 0; o < no; ++o) gb[o] += delta[o];
        for (int i = 0; i < ni; ++i) {
            const float *Wc = W + (int64_t)i * no; double *gc = gW + (int64_t)i * no; float xi = x[i]; float acc = 0.f;
            for (int o = 0; o < no; ++o) { acc += Wc[o] * delta[o]; gc[o] += (double)(delta[o] * xi); }
            dprev[i] = acc;
        }
        for (int i = 0; i < ni; ++i) delta[i] = dprev[i];
    }
    if (n->conv_width) {
        int k = n->conv_width, n0 = n->n_in[0]; const float *w = th + n->conv_off; double *gw = g + n->conv_off;
        for (int i = 0; i < n0; ++i) {
            float d = ab->pre0[i] > 0.f ? delta[i] : 0.f;
            gw[k] += d;
            for (int j = 0; j < k; ++j) { gw[j] += (double)(d * T[i + (k - 1 - j)]); tbar[i + (k - 1 - j)] += w[j] * d; }
        }
    } else {
        for (int i = 0; i < N; ++i) tbar[i] += delta[i];
    }
    for (int k = 0; k < N; ++k) ubar[k] += tbar[k];
}

static float rmsf(const float *x, int n) { float s = 0.f; for (int i = 0; i < n; ++i) s += x[i] * x[i]; return sqrtf(s / (float)n); }

static float initial_dt(const net_t *n, const float *th, const float *u0, const float *f0, const float *colp, float pref,
                        float reltol, float abstol, float dtmax, actbuf_t *ab) {
    int N = n->Nz; float a[MAXW], bb[MAXW], u1[MAXW], f1[MAXW];
    for (int i = 0; i < N; ++i) { float sk = abstol + fabsf(u0[i]) * reltol; a[i] = u0[i] / sk; bb[i] = f0[i] / sk; }
    float d0 = rmsf(a, N), d1 = rmsf(bb, N);
    float dt0 = (d0 < 1e-5f || d1 < 1e-5f) ? 1e-6f : 0.01f * d0 / d1;
    dt0 = fminf(dt0, dtmax);
    for (int i = 0; i < N; ++i) u1[i] = u0[i] + dt0 * f0[i];
    rhs_eval(n, th, u1, colp, pref, f1, ab);
    for (int i = 0; i < N; ++i) { float sk = abstol + fabsf(u0[i]) * reltol; a[i] = (f1[i] - f0[i]) / sk; }
    float d2 = rmsf(a, N) / dt0;
    float m = fmaxf(d1, d2);
    float dt1 = (m <= 1e-15f) ? fmaxf(1e-6f, dt0 * 1e-3f) : powf(10.0f, -(2.0f + log10f(m)) / 5.0f);
    return fminf(fminf(100.0f * dt0, dt1), dtmax);
}

typedef struct { float t, h; int s0, s1; float u[64]; } tape_entry_t; /* Nz <= 64 */

typedef struct { tape_entry_t *e; int n, cap; } tape_t;

static void tape_push(tape_t *tp, float t, float h, int s0, int s1, const float *u, int N) {
    if (tp->n == tp->cap) { tp->cap = tp->cap ? 2 * tp->cap : 256; tp->e = (tape_entry_t *)realloc(tp->e, sizeof(tape_entry_t) * tp->cap); }
    tape_entry_t *e = &tp->e[tp->n++]; e->t = t; e->h = h; e->s0 = s0; e->s1 = s1; memcpy(e->u, u, sizeof(float) * N);
}

/* One column forward.  Tout [n_save][Nz] may be NULL; Ttrue [n_save][Nz] may be NULL; returns sum of squared residuals */
static double solve_column(const net_t *n, const float *th, const float *T0, const float *colp, float pref, const float *saveat, int n_save,
                           float reltol, float abstol, float fixed_dt, float *Tout, const float *Ttrue, int *nacc, int *nrej, int *status,
                           tape_t *tape, actbuf_t *ab, float *k1_final) {
    int N = n->Nz; float u[MAXW], k[7][MAXW], y[MAXW], unew[MAXW];
    double sse = 0.0;
    for (int i = 0; i < N; ++i) u[i] = T0[i];
    float tf = saveat[n_save - 1], t = 0.f; int si = 0; *nacc = 0; *nrej = 0; *status = 0;
    rhs_eval(n, th, u, colp, pref, k[0], ab);
    float dt = fixed_dt > 0.f ? fixed_dt : initial_dt(n, th, u, k[0], colp, pref, reltol, abstol, tf, ab);
    float qold = 1e-4f;
    while (si < n_save && saveat[si] <= t) {
        if (Tout) memcpy(Tout + (int64_t)si * N, u, sizeof(float) * N);
        if (Ttrue) for (int i = 0; i < N; ++i) { double d = (double)u[i] - Ttrue[(int64_t)si * N + i]; sse += d * d; }
        ++si;
    }
    for (int iter = 0; iter < 100000 && t < tf; ++iter) {
        float h = fminf(dt, tf - t);
        int last = (t + h) >= tf * (1.0f - 1e-6f);
        if (last) h = tf - t;
        for (int s = 2; s <= 7; ++s) {
            float a[6]; a_row(s, a);
            for (int i = 0; i < N; ++i) { float acc = 0.f; for (int j = 0; j < s - 1; ++j) acc += a[j] * k[j][i]; y[i] = u[i] + h * acc; }
            if (s == 7) memcpy(unew, y, sizeof(float) * N);
            rhs_eval(n, th, y, colp, pref, k[s - 1], ab);
        }
        float ss = 0.f; int bad = 0;
        for (int i = 0; i < N; ++i) {
            float e = 0.f; for (int j = 0; j < 7; ++j) e += BT[j] * k[j][i];
            e *= h; float sc = abstol + fmaxf(fabsf(u[i]), fabsf(unew[i])) * reltol; e /= sc; ss += e * e;
        }
        float EEst = sqrtf(ss / (float)N);
        if (!isfinite(EEst)) bad = 1;
        int accept; float dtnew;
        if (fixed_dt > 0.f) { accept = 1; dtnew = h; }
        else {
            float q11 = powf(EEst, 0.14f), q = q11 / powf(qold, 0.08f);
            q = fmaxf(0.1f, fminf(5.0f, q / 0.9f));
            accept = EEst <= 1.0f;
            dtnew = accept ? h / q : h / fminf(5.0f, q11 / 0.9f);
        }
        if (bad) { *status = 2; break; }
        if (accept) {
            float tn = last ? tf : t + h; int s0 = si;
            while (si < n_save && saveat[si] <= tn) {
                float th1 = fminf((saveat[si] - t) / h, 1.0f), th2 = th1 * th1, th3 = th2 * th1, th4 = th2 * th2;
                float bth[7]; for (int j = 0; j < 7; ++j) bth[j] = RI[j][0] * th1 + RI[j][1] * th2 + RI[j][2] * th3 + RI[j][3] * th4;
                for (int i = 0; i < N; ++i) {
                    float inc = 0.f; for (int j = 0; j < 7; ++j) inc += bth[j] * k[j][i];
                    float v = u[i] + h * inc;
                    if (Tout) Tout[(int64_t)si * N + i] = v;
                    if (Ttrue) { double d = (double)v - Ttrue[(int64_t)si * N + i]; sse += d * d; }
                }
                ++si;
            }
            if (tape) tape_push(tape, t, h, s0, si, u, N);
            memcpy(u, unew, sizeof(float) * N); memcpy(k[0], k[6], sizeof(float) * N);
            t = tn; if (fixed_dt <= 0.f) qold = fmaxf(EEst, 1e-4f); ++*nacc;
        } else ++*nrej;
        dt = dtnew;
        if (dt < 1e-12f) { *status = 3; break; }
    }
    if (*status == 0 && t < tf) *status = 1;
    if (k1_final) memcpy(k1_final, k[0], sizeof(float) * N);
    return sse;
}

/* Discrete adjoint of one column's accepted steps.  scale = 2/(B*n_save*Nz). */
static void adjoint_column(const net_t *n, const float *th, const float *T0, const float *colp, float pref, const float *saveat, int n_save,
                           const float *Ttrue, const tape_t *tape, double scale, double *g, float *dT0, actbuf_t *ab) {
    int N = n->Nz; float lam[MAXW], kcarry[MAXW];
    float k[7][MAXW], Y[7][MAXW], kb[7][MAXW], lam_n[MAXW], yb[MAXW];
    for (int i = 0; i < N; ++i) { lam[i] = 0.f; kcarry[i] = 0.f; }
    for (int st = tape->n - 1; st >= 0; --st) {
        const tape_entry_t *e = &tape->e[st]; const float *u = e->u; float h = e->h, t = e->t;
        memcpy(Y[0], u, sizeof(float) * N);
        rhs_eval(n, th, u, colp, pref, k[0], ab);
        for (int s = 2; s <= 7; ++s) {
            float a[6]; a_row(s, a);
            for (int i = 0; i < N; ++i) { float acc = 0.f; for (int j = 0; j < s - 1; ++j) acc += a[j] * k[j][i]; Y[s - 1][i] = u[i] + h * acc; }
            rhs_eval(n, th, Y[s - 1], colp, pref, k[s - 1], ab);
        }
        for (int j = 0; j < 7; ++j) for (int i = 0; i < N; ++i) kb[j][i] = 0.f;
        for (int i = 0; i < N; ++i) { kb[6][i] = kcarry[i]; lam_n[i] = 0.f; }
        for (int si = e->s0; si < e->s1; ++si) {
            float th1 = fminf((saveat[si] - t) / h, 1.0f), th2 = th1 * th1, th3 = th2 * th1, th4 = th2 * th2;
            float bth[7]; for (int j = 0; j < 7; ++j) bth[j] = RI[j][0] * th1 + RI[j][1] * th2 + RI[j][2] * th3 + RI[j][3] * th4;
            for (int i = 0; i < N; ++i) {
                float inc = 0.f; for (int j = 0; j < 7; ++j) inc += bth[j] * k[j][i];
                float v = u[i] + h * inc;
                float r = (float)(scale * ((double)v - Ttrue[(int64_t)si * N + i]));
                lam_n[i] += r;
                for (int j = 0; j < 7; ++j) kb[j][i] += h * bth[j] * r;
            }
        }
        for (int s = 7; s >= 2; --s) {
            for (int i = 0; i < N; ++i) yb[i] = (s == 7) ? lam[i] : 0.f;
            rhs_vjp(n, th, Y[s - 1], colp, pref, kb[s - 1], yb, g, ab);
            float a[6]; a_row(s, a);
            for (int i = 0; i < N; ++i) { lam_n[i] += yb[i]; for (int j = 0; j < s - 1; ++j) kb[j][i] += h * a[j] * yb[i]; }
        }
        for (int i = 0; i < N; ++i) { lam[i] = lam_n[i]; kcarry[i] = kb[0][i]; }
    }
    /* saves at t0 and the carried k1-adjoint of the first step */
    for (int si = 0; si < n_save && saveat[si] <= 0.f; ++si)
        for (int i = 0; i < N; ++i) lam[i] += (float)(scale * ((double)T0[i] - Ttrue[(int64_t)si * N + i]));
    rhs_vjp(n, th, T0, colp, pref, kcarry, lam, g, ab);
    if (dT0) memcpy(dT0, lam, sizeof(float) * N);
}

/* ------------------------------------------------------------------------------------------------
 * Continuous adjoint (what DiffEqSensitivity's InterpolatingAdjoint does for src/training.jl:70, restated):
 *   forward pass keeps a dense tape (per accepted step: t, h, u_n and the Tsit5 interpolant in power basis),
 *   backward pass integrates  d(lam)/dt = -J_f(u(t))^T lam  from tf to 0 with its OWN adaptive Tsit5
 *   (same PI controller), u(t) from the tape, lam jumps by dL/du(t_i) at the save times (tstops),
 *   dL/dtheta = integral of (df/dtheta)^T lam dt accumulated with the step's quadrature weights h*b_s.
 * Differences to the reference, on purpose (DESIGN.md "adjoint"): error control acts on lam only (the
 * reference also puts the 24 742 parameter-gradient components into the error norm), and lam is carried
 * un-normalised (jumps = residuals, the factor 2/N applied at the end) so that the tolerance does not
 * depend on the batch size.  No FSAL in the backward pass (every stage contributes to the quadrature).
 * ------------------------------------------------------------------------------------------------ */
typedef struct { float t, h; float u[64], c1[64], c2[64], c3[64], c4[64]; } dtape_entry_t;
typedef struct { dtape_entry_t *e; int n, cap; } dtape_t;

static const float TS_C[7] = {0.f, 0.161f, 0.327f, 0.9f, 0.9800255409045097f, 1.f, 1.f};

static void dtape_push(dtape_t *tp, float t, float h, const float *u, float k[7][MAXW], int N) {
    if (tp->n == tp->cap) { tp->cap = tp->cap ? 2 * tp->cap : 256; tp->e = (dtape_entry_t *)realloc(tp->e, sizeof(dtape_entry_t) * tp->cap); }
    dtape_entry_t *e = &tp->e[tp->n++]; e->t = t; e->h = h;
    for (int i = 0; i < N; ++i) {
        float c1 = 0.f, c2 = 0.f, c3 = 0.f, c4 = 0.f;
        for (int j = 0; j < 7; ++j) { c1 += RI[j][0] * k[j][i]; c2 += RI[j][1] * k[j][i]; c3 += RI[j][2] * k[j][i]; c4 += RI[j][3] * k[j][i]; }
        e->u[i] = u[i]; e->c1[i] = c1; e->c2[i] = c2; e->c3[i] = c3; e->c4[i] = c4;
    }
}

/* forward pass with dense tape + residuals at the save points; returns sum of squared residuals */
static double forward_record(const net_t *n, const float *th, const float *T0, const float *colp, float pref, const float *saveat, int n_save,
                             float reltol, float abstol, float fixed_dt, const float *Ttrue, float *resid, dtape_t *tape, int *nacc, int *nrej,
                             int *status, actbuf_t *ab) {
    int N = n->Nz; float u[MAXW], k[7][MAXW], y[MAXW], unew[MAXW];
    double sse = 0.0;
    for (int i = 0; i < N; ++i) u[i] = T0[i];
    float tf = saveat[n_save - 1], t = 0.f; int si = 0; *nacc = 0; *nrej = 0; *status = 0;
    rhs_eval(n, th, u, colp, pref, k[0], ab);
    float dt = fixed_dt > 0.f ? fixed_dt : initial_dt(n, th, u, k[0], colp, pref, reltol, abstol, tf, ab);
    float qold = 1e-4f;
    while (si < n_save && saveat[si] <= t) {
        for (int i = 0; i < N; ++i) { float r = u[i] - Ttrue[(int64_t)si * N + i]; resid[(int64_t)si * N + i] = r; sse += (double)r * r; }
        ++si;
    }
    for (int iter = 0; iter < 100000 && t < tf; ++iter) {
        float h = fminf(dt, tf - t);
        int last = (t + h) >= tf * (1.0f - 1e-6f);
        if (last) h = tf - t;
        for (int s = 2; s <= 7; ++s) {
            float a[6]; a_row(s, a);
            for (int i = 0; i < N; ++i) { float acc = 0.f; for (int j = 0; j < s - 1; ++j) acc += a[j] * k[j][i]; y[i] = u[i] + h * acc; }
            if (s == 7) memcpy(unew, y, sizeof(float) * N);
            rhs_eval(n, th, y, colp, pref, k[s - 1], ab);
        }
        float ss = 0.f;
        for (int i = 0; i < N; ++i) {
            float e = 0.f; for (int j = 0; j < 7; ++j) e += BT[j] * k[j][i];
            e *= h; float sc = abstol + fmaxf(fabsf(u[i]), fabsf(unew[i])) * reltol; e /= sc; ss += e * e;
        }
        float EEst = sqrtf(ss / (float)N);
        if (!isfinite(EEst)) { *status = 2; break; }
        int accept; float dtnew;
        if (fixed_dt > 0.f) { accept = 1; dtnew = h; }
        else {
            float q11 = powf(EEst, 0.14f), q = q11 / powf(qold, 0.08f);
            q = fmaxf(0.1f, fminf(5.0f, q / 0.9f));
            accept = EEst <= 1.0f;
            dtnew = accept ? h / q : h / fminf(5.0f, q11 / 0.9f);
        }
        if (accept) {
            float tn = last ? tf : t + h;
            dtape_push(tape, t, h, u, k, N);
            const dtape_entry_t *e = &tape->e[tape->n - 1];
            while (si < n_save && saveat[si] <= tn) {
                float th1 = fminf((saveat[si] - t) / h, 1.0f);
                for (int i = 0; i < N; ++i) {
                    float v = e->u[i] + h * (th1 * (e->c1[i] + th1 * (e->c2[i] + th1 * (e->c3[i] + th1 * e->c4[i]))));
                    float r = v - Ttrue[(int64_t)si * N + i]; resid[(int64_t)si * N + i] = r; sse += (double)r * r;
                }
                ++si;
            }
            memcpy(u, unew, sizeof(float) * N); memcpy(k[0], k[6], sizeof(float) * N);
            t = tn; if (fixed_dt <= 0.f) qold = fmaxf(EEst, 1e-4f); ++*nacc;
        } else ++*nrej;
        dt = dtnew;
        if (dt < 1e-12f) { *status = 3; break; }
    }
    if (*status == 0 && t < tf) *status = 1;
    return sse;
}

/* u(t) from the tape; *pn is the running interval index (moved monotonically in either direction) */
static void tape_eval(const dtape_t *tp, int *pn, float t, float *u, int N) {
    int n = *pn;
    while (n > 0 && t < tp->e[n].t) --n;
    while (n + 1 < tp->n && t > tp->e[n].t + tp->e[n].h) ++n;
    *pn = n;
    const dtape_entry_t *e = &tp->e[n];
    float th = (t - e->t) / e->h; th = fminf(fmaxf(th, 0.f), 1.f);
    for (int i = 0; i < N; ++i) u[i] = e->u[i] + e->h * (th * (e->c1[i] + th * (e->c2[i] + th * (e->c3[i] + th * e->c4[i]))));
}

/* backward pass of one column: g (double, un-normalised: caller scales by 2/N_total) */
static void adjoint_continuous(const net_t *n, const float *th, const float *colp, float pref, const float *saveat, int n_save, const float *resid,
                               const dtape_t *tape, float reltol, float abstol, float fixed_dt, double *g, double *gstep, double *gnull, float *lam_out,
                               int *nacc, int *nrej, actbuf_t *ab) {
    int N = n->Nz; float lam[MAXW], lams[7][MAXW], kap[7][MAXW], us[7][MAXW], lnew[MAXW], tmp[MAXW];
    *nacc = 0; *nrej = 0;
    for (int i = 0; i < N; ++i) lam[i] = 0.f;
    if (tape->n == 0) { if (lam_out) memcpy(lam_out, lam, sizeof(float) * N); return; }
    int in = tape->n - 1;
    float t = saveat[n_save - 1];
    float dt = fixed_dt > 0.f ? fixed_dt : tape->e[in].h;
    float qold = 1e-4f;
    int rejrow = 0;      /* rejections in a row */
    for (int si = n_save - 1; si >= 1; --si) {
        for (int i = 0; i < N; ++i) lam[i] += resid[(int64_t)si * N + i];
        float tstop = saveat[si - 1];
        for (int iter = 0; iter < 100000 && t > tstop; ++iter) {
            float h = fminf(dt, t - tstop);
            int last = (t - h) <= tstop + 1e-6f * saveat[n_save - 1] && (t - h) - tstop <= 0.05f * fabsf(h);   /* snap only when it stretches the step by < 5 %: no livelock on rejection */
            if (last) h = t - tstop;
            const float dt_prop = dt;
            for (int p = 0; p < n->n_params; ++p) gstep[p] = 0.0;
            for (int s = 1; s <= 7; ++s) {
                float a[6]; if (s > 1) a_row(s, a);
                for (int i = 0; i < N; ++i) { float acc = 0.f; for (int j = 0; j < s - 1; ++j) acc += a[j] * kap[j][i]; lams[s - 1][i] = lam[i] + h * acc; }
                float ts = (s == 7 && last) ? tstop : t - TS_C[s - 1] * h;
                tape_eval(tape, &in, ts, us[s - 1], N);
                /* kappa_s = J^T lam_s ; parameter quadrature weight h*b_s folded into a scaled copy of lam_s */
                for (int i = 0; i < N; ++i) kap[s - 1][i] = 0.f;
                float bw = 0.f; { float b[6]; a_row(7, b); if (s <= 6) bw = h * b[s - 1]; }
                if (bw != 0.f) {
                    for (int i = 0; i < N; ++i) tmp[i] = bw * lams[s - 1][i];
                    float scaled[MAXW]; for (int i = 0; i < N; ++i) scaled[i] = 0.f;
                    rhs_vjp(n, th, us[s - 1], colp, pref, tmp, scaled, gstep, ab);
                    for (int i = 0; i < N; ++i) kap[s - 1][i] = scaled[i] / bw;
                } else {
                    /* stage 7 has quadrature weight 0: J^T lam is needed (error estimate), its parameter part is discarded */
                    rhs_vjp(n, th, us[s - 1], colp, pref, lams[s - 1], kap[s - 1], gnull, ab);
                }
            }
            memcpy(lnew, lams[6], sizeof(float) * N);
            float ss = 0.f;
            for (int i = 0; i < N; ++i) {
                float e = 0.f; for (int j = 0; j < 7; ++j) e += BT[j] * kap[j][i];
                e *= h; float sc = abstol + fmaxf(fabsf(lam[i]), fabsf(lnew[i])) * reltol; e /= sc; ss += e * e;
            }
            float EEst = sqrtf(ss / (float)N);
            int accept; float dtnew;
            if (fixed_dt > 0.f) { accept = 1; dtnew = fixed_dt; }
            else if (!isfinite(EEst)) { accept = 0; dtnew = h * 0.2f; }
            else {
                float q11 = powf(EEst, 0.14f), q = q11 / powf(qold, 0.08f);
                q = fmaxf(0.1f, fminf(5.0f, q / 0.9f));
                accept = EEst <= 1.0f;
                dtnew = accept ? h / q : h / fminf(5.0f, q11 / 0.9f);
            }
            if (accept) {
                for (int p = 0; p < n->n_params; ++p) g[p] += gstep[p];
                memcpy(lam, lnew, sizeof(float) * N);
                t = last ? tstop : t - h;
                /* a step cut short by the save time (h < the controller's proposal) neither enters the error history nor replaces the
                   proposal (growth capped at 10 h): csrc/nfc_device.cuh:adj_accept_update */
                const int cut = fixed_dt <= 0.f && h < dt_prop;
                if (fixed_dt <= 0.f && !cut) qold = fmaxf(EEst, 1e-4f);
                if (cut) dtnew = fmaxf(dtnew, fminf(dt_prop, 10.f * h));
                ++*nacc;
                rejrow = 0;
            } else {
                /* J(u(t))^T jumps in time where a gradient changes sign: there the error falls like h, not h^5; from the second rejection
                   in a row on the new step assumes that (csrc/nfc_device.cuh:adj_reject_dt) */
                if (rejrow >= 1 && isfinite(EEst)) dtnew = fminf(dtnew, fmaxf(0.2f * h, 0.9f * h / EEst));
                ++rejrow;
                ++*nrej;
            }
            dt = dtnew;
            if (dt < 1e-12f) break;
        }
    }
    for (int i = 0; i < N; ++i) lam[i] += resid[i];     /* save point at t0 (only dL/dT0 sees it) */
    if (lam_out) memcpy(lam_out, lam, sizeof(float) * N);
}

int oracle_loss_grad_continuous(int Nz, int conv_width, int n_layers, const int *dims, const int *acts, const float *theta,
                                const float *T0, const float *colp, const float *glob, const float *saveat, int n_save,
                                float reltol, float abstol, float fixed_dt, float adj_reltol, float adj_abstol, const float *Ttrue, double *loss,
                                float *grad, float *dT0, int32_t *adj_naccept, int32_t *adj_nreject, int64_t B, int nthreads) {
    net_t n; if (net_init(&n, Nz, conv_width, n_layers, dims, acts)) return -1;
    float pref = glob[1] / glob[0] * glob[3] / glob[2];
    if (nthreads < 1) nthreads = 1;
    double ntot = (double)B * n_save * Nz, sse_all = 0.0;
    double *gall = (double *)calloc((size_t)n.n_params * nthreads, sizeof(double));
#pragma omp parallel num_threads(nthreads) reduction(+ : sse_all)
    {
        int tid = 0;
#ifdef _OPENMP
        tid = omp_get_thread_num();
#endif
        actbuf_t *ab = (actbuf_t *)malloc(sizeof(actbuf_t));
        double *g = gall + (size_t)tid * n.n_params;
        double *gstep = (double *)malloc(sizeof(double) * n.n_params);
        double *gnull = (double *)malloc(sizeof(double) * n.n_params);
        float *resid = (float *)malloc(sizeof(float) * n_save * Nz);
#pragma omp for schedule(dynamic, 1)
        for (int64_t b = 0; b < B; ++b) {
            int na, nr, st, ba, br; dtape_t tape = {0, 0, 0}; float lam0[MAXW];
            sse_all += forward_record(&n, theta, T0 + b * Nz, colp + b * 3, pref, saveat, n_save, reltol, abstol, fixed_dt,
                                      Ttrue + b * (int64_t)n_save * Nz, resid, &tape, &na, &nr, &st, ab);
            adjoint_continuous(&n, theta, colp + b * 3, pref, saveat, n_save, resid, &tape, adj_reltol, adj_abstol, fixed_dt, g, gstep, gnull, lam0, &ba, &br, ab);
            if (dT0) for (int i = 0; i < Nz; ++i) dT0[b * Nz + i] = (float)(2.0 / ntot * lam0[i]);
            if (adj_naccept) adj_naccept[b] = ba; if (adj_nreject) adj_nreject[b] = br;
            free(tape.e);
        }
        free(ab); free(gstep); free(gnull); free(resid);
    }
    for (int64_t p = 0; p < n.n_params; ++p) { double s = 0.0; for (int t = 0; t < nthreads; ++t) s += gall[(size_t)t * n.n_params + p]; grad[p] = (float)(2.0 / ntot * s); }
    free(gall);
    *loss = sse_all / ntot;
    return 0;
}

/* ------------------------------- exported entry points ------------------------------- */
int oracle_n_params(int Nz, int conv_width, int n_layers, const int *dims, const int *acts, int64_t *out) {
    net_t n; if (net_init(&n, Nz, conv_width, n_layers, dims, acts)) return -1; *out = n.n_params; return 0;
}

int oracle_rhs(int Nz, int conv_width, int n_layers, const int *dims, const int *acts, const float *theta,
               const float *T, const float *colp, const float *glob, float *out, int64_t B, int nthreads) {
    net_t n; if (net_init(&n, Nz, conv_width, n_layers, dims, acts)) return -1;
    float pref = glob[1] / glob[0] * glob[3] / glob[2];
    if (nthreads < 1) nthreads = 1;
#pragma omp parallel num_threads(nthreads)
    {
        actbuf_t *ab = (actbuf_t *)malloc(sizeof(actbuf_t));
#pragma omp for schedule(dynamic, 64)
        for (int64_t b = 0; b < B; ++b) rhs_eval(&n, theta, T + b * Nz, colp + b * 3, pref, out + b * Nz, ab);
        free(ab);
    }
    return 0;
}

int oracle_solve(int Nz, int conv_width, int n_layers, const int *dims, const int *acts, const float *theta,
                 const float *T0, const float *colp, const float *glob, const float *saveat, int n_save,
                 float reltol, float abstol, float fixed_dt, float *Tout, int32_t *naccept, int32_t *nreject, int32_t *status,
                 int64_t B, int nthreads) {
    net_t n; if (net_init(&n, Nz, conv_width, n_layers, dims, acts)) return -1;
    float pref = glob[1] / glob[0] * glob[3] / glob[2];
    if (nthreads < 1) nthreads = 1;
#pragma omp parallel num_threads(nthreads)
    {
        actbuf_t *ab = (actbuf_t *)malloc(sizeof(actbuf_t));
#pragma omp for schedule(dynamic, 4)
        for (int64_t b = 0; b < B; ++b) {
            int na, nr, st;
            solve_column(&n, theta, T0 + b * Nz, colp + b * 3, pref, saveat, n_save, reltol, abstol, fixed_dt,
                         Tout ? Tout + b * (int64_t)n_save * Nz : NULL, NULL, &na, &nr, &st, NULL, ab, NULL);
            if (naccept) naccept[b] = na; if (nreject) nreject[b] = nr; if (status) status[b] = st;
        }
        free(ab);
    }
    return 0;
}

int oracle_loss_grad(int Nz, int conv_width, int n_layers, const int *dims, const int *acts, const float *theta,
                     const float *T0, const float *colp, const float *glob, const float *saveat, int n_save,
                     float reltol, float abstol, float fixed_dt, const float *Ttrue, double *loss, float *grad, float *dT0,
                     int32_t *naccept, int64_t B, int nthreads) {
    net_t n; if (net_init(&n, Nz, conv_width, n_layers, dims, acts)) return -1;
    float pref = glob[1] / glob[0] * glob[3] / glob[2];
    if (nthreads < 1) nthreads = 1;
    double ntot = (double)B * n_save * Nz, scale = 2.0 / ntot, sse_all = 0.0;
    double *gall = (double *)calloc((size_t)n.n_params * nthreads, sizeof(double));
#pragma omp parallel num_threads(nthreads) reduction(+ : sse_all)
    {
        int tid = 0;
#ifdef _OPENMP
        tid = omp_get_thread_num();
#endif
        actbuf_t *ab = (actbuf_t *)malloc(sizeof(actbuf_t));
        double *g = gall + (size_t)tid * n.n_params;
#pragma omp for schedule(dynamic, 1)
        for (int64_t b = 0; b < B; ++b) {
            int na, nr, st; tape_t tape = {0, 0, 0};
            sse_all += solve_column(&n, theta, T0 + b * Nz, colp + b * 3, pref, saveat, n_save, reltol, abstol, fixed_dt, NULL,
                                    Ttrue + b * (int64_t)n_save * Nz, &na, &nr, &st, &tape, ab, NULL);
            adjoint_column(&n, theta, T0 + b * Nz, colp + b * 3, pref, saveat, n_save, Ttrue + b * (int64_t)n_save * Nz, &tape, scale, g,
                           dT0 ? dT0 + b * Nz : NULL, ab);
            if (naccept) naccept[b] = na;
            free(tape.e);
        }
        free(ab);
    }
    for (int64_t p = 0; p < n.n_params; ++p) { double s = 0.0; for (int t = 0; t < nthreads; ++t) s += gall[(size_t)t * n.n_params + p]; grad[p] = (float)s; }
    free(gall);
    *loss = sse_all / ntot;
    return 0;
}

/* test hook: ubar[B][Nz] = J^T kbar, g[n_params] += dtheta contributions (single thread) */
int oracle_rhs_vjp(int Nz, int conv_width, int n_layers, const int *dims, const int *acts, const float *theta,
                   const float *T, const float *colp, const float *glob, const float *kbar, float *ubar, double *g, int64_t B) {
    net_t n; if (net_init(&n, Nz, conv_width, n_layers, dims, acts)) return -1;
    float pref = glob[1] / glob[0] * glob[3] / glob[2];
    actbuf_t *ab = (actbuf_t *)malloc(sizeof(actbuf_t));
    for (int64_t b = 0; b < B; ++b) {
        for (int i = 0; i < Nz; ++i) ubar[b * Nz + i] = 0.f;
        rhs_vjp(&n, theta, T + b * Nz, colp + b * 3, pref, kbar + b * Nz, ubar + b * Nz, g, ab);
    }
    free(ab);
    return 0;
}
