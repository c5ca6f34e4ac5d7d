/*
 * mpc_oracle_laned.c — CPU restatement of the SAME problem and solver as
 * mpc_oracle.c, but with the floating-point operation ORDER of the CUDA kernel
 * (dyobav_mpcnwta_warehouse_b200/csrc/mpcb_device.cuh "ARITHMETIC CONTRACT"):
 * horizon sums are 32-lane Kogge-Stone scans and xor-butterfly reductions,
 * products fuse exactly where the kernel calls fma(), sin/cos come from the
 * same Cody-Waite + fdlibm-polynomial routine.  TEST INFRASTRUCTURE ONLY.
 *
 * Why it exists: PANOC on this non-smooth problem amplifies round-off (its
 * Lipschitz estimate divides a gradient difference by |h| = 6e-12), so two
 * correct implementations that merely round differently drift apart within a
 * few iterations (tests/test_oracle_sensitivity.py).  With identical rounding
 * the GPU must reproduce this file BIT FOR BIT over whole ALM/PANOC runs, which
 * turns parity into an exact test.  Its own correctness is anchored on
 * mpc_oracle.c: same psi / grad psi to round-off (tests/test_oracle_laned.py),
 * and mpc_oracle.c is pinned to the reference's code by tests/golden.
 *
 * Written from the arithmetic contract, not from the kernel source: a warp is
 * an array of 32 lanes, lane l of row j owns horizon step k = l + 32 j.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../include/mpcb.h"

#define W 32
#define JMAX 2
#define MAXN (W * JMAX)

typedef double lanes_t[JMAX][W];

typedef struct {
    int N, J, Nother, Nstc, nedge, Ndyn;
    int G;                     /* worker groups of team mode (0: one-warp order) */
    double ts, k6, inv_ts, ds2, vmargin, smargin;
    double vmin, vmax, wmax, amin, amax, wamax;
    /* staged scenario */
    double s0[3], um1[2], sN[3], q[10];
    double rv[MAXN], qstc[MAXN];
    double sgx[MAXN], sgy[MAXN], sdx[MAXN], sdy[MAXN], sinv[MAXN];
    double *c0x, *c0y;         /* [Nother] */
    double *cx, *cy;           /* [Nother][N] */
    double* poly;              /* [Nstc][nedge][3] = b, -a0, -a1 */
    double* e0;                /* [9][Ndyn] */
    double* et;                /* [9][Ndyn][N] */
} scen_t;

enum { E_CX = 0, E_CY, E_CA, E_SA, E_I1I, E_I2I, E_I1R, E_I2R, E_WAL, EF };

/* ---- sin/cos: Cody-Waite by pi/2 + fdlibm kernel polynomials (Horner, fma) ---- */
/* max(0, r) and clamp to [0, 1] as the kernel writes them (plain selects; NaN -> 0) */
static inline double pos_part(double r) { return r > 0.0 ? r : 0.0; }
static inline double clamp01(double t) { return t > 0.0 ? (t < 1.0 ? t : 1.0) : 0.0; }

static void sincos_cw(double x, double* sn, double* cs)
{
    const double fn = rint(x * 6.36619772367581382433e-01);
    double r = fma(-fn, 1.57079632673412561417e+00, x);
    r = fma(-fn, 6.07710050650619224932e-11, r);
    const double z = r * r;
    double ps = fma(z, 1.58969099521155010221e-10, -2.50507602534068634195e-08);
    ps = fma(z, ps, 2.75573137070700676789e-06);
    ps = fma(z, ps, -1.98412698298579493134e-04);
    ps = fma(z, ps, 8.33333333332248946124e-03);
    ps = fma(z, ps, -1.66666666666666324348e-01);
    const double s = fma(r * z, ps, r);
    double pc = fma(z, -1.13596475577881948265e-11, 2.08757232129817482790e-09);
    pc = fma(z, pc, -2.75573143513906633035e-07);
    pc = fma(z, pc, 2.48015872894767294178e-05);
    pc = fma(z, pc, -1.38888888888741095749e-03);
    pc = fma(z, pc, 4.16666666666666019037e-02);
    const double c = fma(z * z, pc, fma(-0.5, z, 1.0));
    const int q = (int)fn & 3;
    const double s1 = (q & 1) ? c : s;
    const double c1 = (q & 1) ? s : c;
    *sn = (q & 2) ? -s1 : s1;
    *cs = ((q + 1) & 2) ? -c1 : c1;
}

/* ---- 32-lane collectives ---- */
static double warp_sum(const double in[W])
{
    double v[W], t[W];
    memcpy(v, in, sizeof(v));
    for (int m = 16; m > 0; m >>= 1) {
        for (int l = 0; l < W; ++l) t[l] = v[l] + v[l ^ m];
        memcpy(v, t, sizeof(v));
    }
    return v[0];
}
static void scan_incl(double v[W])
{
    double t[W];
    for (int d = 1; d < W; d <<= 1) {
        for (int l = 0; l < W; ++l) t[l] = l >= d ? v[l] + v[l - d] : v[l];
        memcpy(v, t, sizeof(t));
    }
}
static void rscan_incl(double v[W])
{
    double t[W];
    for (int d = 1; d < W; d <<= 1) {
        for (int l = 0; l < W; ++l) t[l] = l + d < W ? v[l] + v[l + d] : v[l];
        memcpy(v, t, sizeof(t));
    }
}
static void prefix(int J, lanes_t a, lanes_t excl, lanes_t incl)
{
    double carry = 0.0;
    for (int j = 0; j < J; ++j) {
        double s[W];
        memcpy(s, a[j], sizeof(s));
        scan_incl(s);
        for (int l = 0; l < W; ++l) {
            double e = l == 0 ? 0.0 : s[l - 1];
            if (excl) excl[j][l] = carry + e;
            incl[j][l] = carry + s[l];
        }
        if (J > 1) carry += s[W - 1];
    }
}
static void suffix(int J, lanes_t a, lanes_t excl, lanes_t incl)
{
    double carry = 0.0;
    for (int j = J - 1; j >= 0; --j) {
        double s[W];
        memcpy(s, a[j], sizeof(s));
        rscan_incl(s);
        for (int l = 0; l < W; ++l) {
            double e = l == W - 1 ? 0.0 : s[l + 1];
            if (excl) excl[j][l] = carry + e;
            if (incl) incl[j][l] = carry + s[l];
        }
        if (J > 1) carry += s[0];
    }
}
static double sumsq2(int J, lanes_t a, lanes_t b)
{
    double p[W];
    for (int l = 0; l < W; ++l) {
        double s = 0.0;
        for (int j = 0; j < J; ++j) s = fma(a[j][l], a[j][l], fma(b[j][l], b[j][l], s));
        p[l] = s;
    }
    return warp_sum(p);
}
static double dotw(int J, lanes_t a0, lanes_t a1, lanes_t b0, lanes_t b1)
{
    double p[W];
    for (int l = 0; l < W; ++l) {
        double s = 0.0;
        for (int j = 0; j < J; ++j) s = fma(a0[j][l], b0[j][l], fma(a1[j][l], b1[j][l], s));
        p[l] = s;
    }
    return warp_sum(p);
}

/* Team mode (arithmetic contract, csrc/mpcb_device.cuh "team mode"): dimension sets with at
 * least 64 ellipses are solved by 12-warp CTAs whose 10 worker warps own (step, group) pairs;
 * the number of groups is the largest power of two G <= 32 with G*N <= 320.  The polygon and
 * ellipse terms of a step are summed per group i % G (polygons, then ellipses, index order, from
 * +0.0), the group sums are added together in group order (from +0.0) and the total is added to
 * the step's accumulators; the F2 share of the gradient adds both slots of an ellipse before the
 * fma. */
static __thread int g_force_team = 0;     /* cfg->team_mode == 1 (set per call), or mpcl_set_team_mode for mpcl_eval */
static int g_force_team_global = 0;
void mpcl_set_team_mode(int on) { g_force_team_global = on; }

static int team_groups_for(const mpcb_dims* d, int force)
{
    if (d->Ndyn < 64 && !force) return 0;
    int g = 1;
    while (2 * g <= 32 && 2 * g * d->N <= 32 * 10) g *= 2;
    return g;
}
int32_t mpcl_team_groups_cfg(const mpcb_dims* d, const mpcb_solver_cfg* c) { return team_groups_for(d, c && c->team_mode == 1); }

int32_t mpcl_team_groups(const mpcb_dims* d)
{
    if (d->Ndyn < 64) return 0;
    int g = 1;
    while (2 * g <= 32 && 2 * g * d->N <= 32 * 10) g *= 2;
    return g;
}

/* ---- staging (K3): raw parameter row -> scenario ---- */
static void scen_free(scen_t* S)
{
    free(S->c0x); free(S->cx); free(S->poly); free(S->e0); free(S->et);
}
static int scen_stage(scen_t* S, const mpcb_dims* d, const mpcb_robot* rb, const double* p)
{
    memset(S, 0, sizeof(*S));
    const int N = d->N;
    if (N < 1 || N > MAXN || d->nedge < 1 || d->nedge > MPCB_MAX_EDGE) return MPCB_E_DIMS;
    S->N = N; S->J = N <= W ? 1 : 2;
    S->Nother = d->Nother; S->Nstc = d->Nstc; S->nedge = d->nedge; S->Ndyn = d->Ndyn;
    S->G = team_groups_for(d, g_force_team || g_force_team_global);
    S->ts = rb->ts; S->k6 = rb->ts / 6.0; S->inv_ts = 1.0 / rb->ts;
    S->ds2 = rb->vehicle_width * rb->vehicle_width;
    S->vmargin = rb->vehicle_margin; S->smargin = rb->social_margin;
    S->vmin = rb->lin_vel_min; S->vmax = rb->lin_vel_max; S->wmax = rb->ang_vel_max;
    S->amin = rb->lin_acc_min; S->amax = rb->lin_acc_max; S->wamax = rb->ang_acc_max;
    int o = 0;
    const double* um1 = p + o; o += 2;
    const double* s0 = p + o; o += 3;
    const double* sN = p + o; o += 3;
    const double* q = p + o; o += 10;
    const double* rs = p + o; o += 3 * N;
    const double* rv = p + o; o += N;
    const double* c0 = p + o; o += 3 * d->Nother;
    const double* c = p + o; o += 3 * N * d->Nother;
    const double* os = p + o; o += 3 * d->nedge * d->Nstc;
    const double* od = p + o; o += 6 * (N + 1) * d->Ndyn;
    const double* qstc = p + o; o += N;
    const double* qdyn = p + o;
    memcpy(S->um1, um1, 16); memcpy(S->s0, s0, 24); memcpy(S->sN, sN, 24); memcpy(S->q, q, 80);
    for (int k = 0; k < N; ++k) {
        S->rv[k] = rv[k];
        S->qstc[k] = qstc[k];
        const int k2 = k + 1 < N ? k + 1 : N - 1;
        const double ax = rs[3 * k], ay = rs[3 * k + 1];
        const double dx = rs[3 * k2] - ax, dy = rs[3 * k2 + 1] - ay;
        S->sgx[k] = ax; S->sgy[k] = ay; S->sdx[k] = dx; S->sdy[k] = dy;
        S->sinv[k] = 1.0 / (dx * dx + dy * dy + 1e-16);
    }
    S->c0x = (double*)calloc((size_t)(2 * d->Nother + 1), 8); S->c0y = S->c0x + d->Nother;
    S->cx = (double*)calloc((size_t)(2 * d->Nother * N + 1), 8); S->cy = S->cx + d->Nother * N;
    for (int r = 0; r < d->Nother; ++r) {
        S->c0x[r] = c0[3 * r]; S->c0y[r] = c0[3 * r + 1];
        for (int k = 0; k < N; ++k) {
            S->cx[r * N + k] = c[r * 3 * N + 3 * k];
            S->cy[r * N + k] = c[r * 3 * N + 3 * k + 1];
        }
    }
    S->poly = (double*)calloc((size_t)(3 * d->nedge * d->Nstc + 1), 8);
    for (int i = 0; i < d->Nstc; ++i)
        for (int e = 0; e < d->nedge; ++e) {
            const double* qq = os + i * 3 * d->nedge;
            S->poly[(i * d->nedge + e) * 3] = qq[e];
            S->poly[(i * d->nedge + e) * 3 + 1] = -qq[d->nedge + e];
            S->poly[(i * d->nedge + e) * 3 + 2] = -qq[2 * d->nedge + e];
        }
    S->e0 = (double*)calloc((size_t)(EF * d->Ndyn + 1), 8);
    S->et = (double*)calloc((size_t)(EF * d->Ndyn * N + 1), 8);
    for (int ob = 0; ob < d->Ndyn; ++ob)
        for (int t = 0; t <= N; ++t) {
            const double* qq = od + (size_t)(ob * (N + 1) + t) * 6;
            const double rx = qq[2], ry = qq[3];
            const double rxi = t == 0 ? rx + S->vmargin + S->smargin : rx + S->vmargin;
            const double ryi = t == 0 ? ry + S->vmargin + S->smargin : ry + S->vmargin;
            const double wgt = t == 0 ? 1000.0 : qdyn[t - 1];
            double f[EF], sa, ca;
            sincos_cw(qq[4], &sa, &ca);
            f[E_CX] = qq[0]; f[E_CY] = qq[1]; f[E_CA] = ca; f[E_SA] = sa;
            f[E_I1I] = 1.0 / ((rxi + 1e-6) * (rxi + 1e-6));
            f[E_I2I] = 1.0 / ((ryi + 1e-6) * (ryi + 1e-6));
            f[E_I1R] = 1.0 / ((rx + 1e-6) * (rx + 1e-6));
            f[E_I2R] = 1.0 / ((ry + 1e-6) * (ry + 1e-6));
            f[E_WAL] = wgt * qq[5];
            for (int m = 0; m < EF; ++m) {
                if (t == 0) S->e0[m * d->Ndyn + ob] = f[m];
                else S->et[(m * d->Ndyn + ob) * N + (t - 1)] = f[m];
            }
        }
    return MPCB_OK;
}

typedef struct { double cost, gx, gy, hr, hrx, hry; } ell_t;

static void ellipse_terms(int GRAD, const double* f, int stride, double x, double y, ell_t* o)
{
    const double ex = x - f[E_CX * stride], ey = y - f[E_CY * stride];
    const double ca = f[E_CA * stride], sa = f[E_SA * stride];
    const double a = fma(ex, ca, ey * sa);
    const double b = fma(ex, sa, -(ey * ca));
    const double a2 = a * a, b2 = b * b;
    const double i1 = f[E_I1I * stride], i2 = f[E_I2I * stride];
    const double Ei = fma(-b2, i2, fma(-a2, i1, 1.0));
    o->cost = 0.0; o->gx = 0.0; o->gy = 0.0; o->hr = 0.0; o->hrx = 0.0; o->hry = 0.0;
    if (Ei > 0.0) {
        const double wal = f[E_WAL * stride];
        o->cost = wal * (Ei * Ei);
        if (GRAD) {
            const double ta = 2.0 * a * i1, tb = 2.0 * b * i2;
            const double dEx = -fma(ta, ca, tb * sa);
            const double dEy = -fma(ta, sa, -(tb * ca));
            const double m = 2.0 * wal * Ei;
            o->gx = m * dEx;
            o->gy = m * dEy;
        }
        const double r1 = f[E_I1R * stride], r2 = f[E_I2R * stride];
        const double Er = fma(-b2, r2, fma(-a2, r1, 1.0));
        if (Er > 0.0) {
            o->hr = Er;
            if (GRAD) {
                const double ta = 2.0 * a * r1, tb = 2.0 * b * r2;
                o->hrx = -fma(ta, ca, tb * sa);
                o->hry = -fma(ta, sa, -(tb * ca));
            }
        }
    }
}

static double polygon_ind(int GRAD, const double* e, int nedge, double x, double y, double* dIx,
                          double* dIy)
{
    double I = 1.0;
    for (int j = 0; j < nedge; ++j) {
        const double r = fma(e[3 * j + 2], y, fma(e[3 * j + 1], x, e[3 * j]));
        I *= pos_part(r);
    }
    *dIx = 0.0; *dIy = 0.0;
    if (GRAD && I > 0.0) {
        for (int j = 0; j < nedge; ++j) {
            double pr = 1.0;
            for (int m = 0; m < nedge; ++m)
                if (m != j) pr *= fma(e[3 * m + 2], y, fma(e[3 * m + 1], x, e[3 * m]));
            *dIx = fma(pr, e[3 * j + 1], *dIx);
            *dIy = fma(pr, e[3 * j + 2], *dIy);
        }
    }
    return I;
}

typedef struct { double psi, f, f2sq; lanes_t gv, gw; } eval_out_t;

/* psi(u; c, y) and its gradient, kernel operation order (all terms evaluated) */
static void eval_psi(const scen_t* S, lanes_t v, lanes_t w, double c, lanes_t ya, lanes_t yw,
                     int GRAD, eval_out_t* out, double* F1out, double* F2out)
{
    const int N = S->N, J = S->J;
    const double* q = S->q;
    int act[JMAX][W], kk[JMAX][W];
    lanes_t dth, th, thn, c0s, s0s, cb, sb, cc, sc, Cc, Ss, dx, dy, px, py;
    for (int j = 0; j < J; ++j)
        for (int l = 0; l < W; ++l) {
            kk[j][l] = l + W * j;
            act[j][l] = kk[j][l] < N;
            dth[j][l] = act[j][l] ? S->ts * w[j][l] : 0.0;
        }
    prefix(J, dth, th, thn);
    for (int j = 0; j < J; ++j)
        for (int l = 0; l < W; ++l) {
            const double t0 = S->s0[2] + th[j][l];
            const double tb = t0 + 0.5 * dth[j][l];
            sincos_cw(t0, &s0s[j][l], &c0s[j][l]);
            sincos_cw(tb, &sb[j][l], &cb[j][l]);
        }
    for (int j = 0; j < J; ++j)
        for (int l = 0; l < W; ++l) {
            double cn = l + 1 < W ? c0s[j][l + 1] : c0s[j][l];
            double sn = l + 1 < W ? s0s[j][l + 1] : s0s[j][l];
            if (j + 1 < J) {
                if (l == W - 1) { cn = c0s[j + 1][0]; sn = s0s[j + 1][0]; }
            } else if (N == W * J) {
                if (l == W - 1) sincos_cw(S->s0[2] + thn[j][l], &sn, &cn);
            }
            cc[j][l] = cn; sc[j][l] = sn;
            Cc[j][l] = c0s[j][l] + 4.0 * cb[j][l] + cc[j][l];
            Ss[j][l] = s0s[j][l] + 4.0 * sb[j][l] + sc[j][l];
            const double kv = act[j][l] ? S->k6 * v[j][l] : 0.0;
            dx[j][l] = kv * Cc[j][l];
            dy[j][l] = kv * Ss[j][l];
        }
    prefix(J, dx, 0, px);
    prefix(J, dy, 0, py);

    const double qvel = q[1], rv = q[3], rw = q[4], qrpd = q[7];
    double cost[W];
    lanes_t gx, gy, gvd, gwd, Spoly, dSx, dSy, X, Y;
    for (int l = 0; l < W; ++l) cost[l] = 0.0;
    for (int j = 0; j < J; ++j)
        for (int l = 0; l < W; ++l) {
            const int k = act[j][l] ? kk[j][l] : N - 1;
            const double x = S->s0[0] + px[j][l], y = S->s0[1] + py[j][l];
            X[j][l] = x; Y[j][l] = y;
            double cst = 0.0, ggx = 0.0, ggy = 0.0;
            {   /* reference path */
                double best = INFINITY;
                int ib = k;
                for (int i = k; i < N; ++i) {
                    const double ex = x - S->sgx[i], ey = y - S->sgy[i];
                    const double ddx = S->sdx[i], ddy = S->sdy[i];
                    const double th_ = fma(ex, ddx, ey * ddy) * S->sinv[i];
                    const double ts_ = clamp01(th_);
                    const double vx = fma(ts_, ddx, -ex), vy = fma(ts_, ddy, -ey);
                    const double d2 = fma(vx, vx, vy * vy);
                    if (d2 < best) { best = d2; ib = i; }
                }
                cst = best * qrpd;
                if (GRAD) {
                    const double ex = x - S->sgx[ib], ey = y - S->sgy[ib];
                    const double ddx = S->sdx[ib], ddy = S->sdy[ib], inv = S->sinv[ib];
                    const double th_ = fma(ex, ddx, ey * ddy) * inv;
                    const double ts_ = clamp01(th_);
                    const double vx = fma(ts_, ddx, -ex), vy = fma(ts_, ddy, -ey);
                    const double dt = (th_ > 0.0 && th_ < 1.0) ? 1.0 : ((th_ == 0.0 || th_ == 1.0) ? 0.5 : 0.0);
                    const double vd = fma(vx, ddx, vy * ddy) * dt * inv;
                    ggx = 2.0 * qrpd * fma(vd, ddx, -vx);
                    ggy = 2.0 * qrpd * fma(vd, ddy, -vy);
                }
            }
            {   /* speed reference + control effort */
                const double vv = v[j][l], ww = w[j][l];
                const double dv = vv - S->rv[k];
                cst = fma(qvel, dv * dv, cst);
                cst += rv * (vv * vv) + rw * (ww * ww);
                gvd[j][l] = 2.0 * qvel * dv + 2.0 * rv * vv;
                gwd[j][l] = 2.0 * rw * ww;
            }
            {   /* fleet */
                double s1 = 0.0, s2 = 0.0;
                for (int r = 1; r < S->Nother; ++r) {
                    const double ex = x - S->c0x[r], ey = y - S->c0y[r];
                    const double h = S->ds2 - fma(ex, ex, ey * ey);
                    if (h > 0.0) {
                        s1 += h;
                        if (GRAD) { ggx = fma(-2000.0, ex, ggx); ggy = fma(-2000.0, ey, ggy); }
                    }
                }
                for (int r = 0; r < S->Nother; ++r) {
                    const double ex = x - S->cx[r * N + k], ey = y - S->cy[r * N + k];
                    const double h = S->ds2 - fma(ex, ex, ey * ey);
                    if (h > 0.0) {
                        s2 += h;
                        if (GRAD) { ggx = fma(-20.0, ex, ggx); ggy = fma(-20.0, ey, ggy); }
                    }
                }
                cst = fma(1000.0, s1, cst);
                cst = fma(10.0, s2, cst);
            }
            double sp = 0.0, spx = 0.0, spy = 0.0;
            if (S->G == 0) {   /* static polygons, then dynamic ellipses, one chain per step */
                const double qs = S->qstc[k];
                for (int i = 0; i < S->Nstc; ++i) {
                    double dIx, dIy;
                    const double I = polygon_ind(GRAD, S->poly + i * 3 * S->nedge, S->nedge, x, y, &dIx, &dIy);
                    if (I > 0.0) {
                        cst = fma(qs, I * I, cst);
                        sp += I;
                        if (GRAD) {
                            const double m = 2.0 * qs * I;
                            ggx = fma(m, dIx, ggx);
                            ggy = fma(m, dIy, ggy);
                            spx += dIx;
                            spy += dIy;
                        }
                    }
                }
                for (int i = 0; i < S->Ndyn; ++i) {
                    ell_t a, b;
                    ellipse_terms(GRAD, S->e0 + i, S->Ndyn, x, y, &a);
                    cst += a.cost;
                    if (GRAD) { ggx += a.gx; ggy += a.gy; }
                    ellipse_terms(GRAD, S->et + k + i * N, S->Ndyn * N, x, y, &b);
                    cst += b.cost;
                    if (GRAD) { ggx += b.gx; ggy += b.gy; }
                }
            } else {
                /* team mode: group g sums the polygons i % G == g, then the ellipses i % G == g
                   (index order, from +0.0); the group sums are added in group order */
                const double qs = S->qstc[k];
                double pc[32], pgx[32], pgy[32], psp[32], pspx[32], pspy[32];
                for (int g = 0; g < S->G; ++g) { pc[g] = 0.0; pgx[g] = 0.0; pgy[g] = 0.0; psp[g] = 0.0; pspx[g] = 0.0; pspy[g] = 0.0; }
                for (int i = 0; i < S->Nstc; ++i) {
                    const int g = i % S->G;
                    double dIx, dIy;
                    const double I = polygon_ind(GRAD, S->poly + i * 3 * S->nedge, S->nedge, x, y, &dIx, &dIy);
                    if (I > 0.0) {
                        pc[g] = fma(qs, I * I, pc[g]);
                        psp[g] += I;
                        if (GRAD) {
                            const double m = 2.0 * qs * I;
                            pgx[g] = fma(m, dIx, pgx[g]);
                            pgy[g] = fma(m, dIy, pgy[g]);
                            pspx[g] += dIx;
                            pspy[g] += dIy;
                        }
                    }
                }
                for (int i = 0; i < S->Ndyn; ++i) {
                    const int g = i % S->G;
                    ell_t a, b;
                    ellipse_terms(GRAD, S->e0 + i, S->Ndyn, x, y, &a);
                    pc[g] += a.cost;
                    if (GRAD) { pgx[g] += a.gx; pgy[g] += a.gy; }
                    ellipse_terms(GRAD, S->et + k + i * N, S->Ndyn * N, x, y, &b);
                    pc[g] += b.cost;
                    if (GRAD) { pgx[g] += b.gx; pgy[g] += b.gy; }
                }
                double tc = 0.0, tgx = 0.0, tgy = 0.0;
                for (int g = 0; g < S->G; ++g) {
                    tc += pc[g];
                    sp += psp[g];
                    if (GRAD) { tgx += pgx[g]; tgy += pgy[g]; spx += pspx[g]; spy += pspy[g]; }
                }
                cst += tc;
                if (GRAD) { ggx += tgx; ggy += tgy; }
            }
            if (!act[j][l]) { cst = 0.0; ggx = 0.0; ggy = 0.0; sp = 0.0; spx = 0.0; spy = 0.0; gvd[j][l] = 0.0; gwd[j][l] = 0.0; }
            cost[l] += cst;
            gx[j][l] = ggx; gy[j][l] = ggy;
            Spoly[j][l] = sp; dSx[j][l] = spx; dSy[j][l] = spy;
        }
    /* terminal */
    double gthN[W];
    for (int l = 0; l < W; ++l) gthN[l] = 0.0;
    for (int j = 0; j < J; ++j)
        for (int l = 0; l < W; ++l)
            if (act[j][l] && kk[j][l] == N - 1) {
                const double qN = q[5], qthN = q[6];
                const double ex = S->s0[0] + px[j][l] - S->sN[0], ey = S->s0[1] + py[j][l] - S->sN[1];
                const double et = S->s0[2] + thn[j][l] - S->sN[2];
                cost[l] += qN * fma(ex, ex, ey * ey) + qthN * (et * et);
                gx[j][l] = fma(2.0 * qN, ex, gx[j][l]);
                gy[j][l] = fma(2.0 * qN, ey, gy[j][l]);
                gthN[l] = 2.0 * qthN * et;
            }
    /* penalty constraints F2 */
    double f2sq = 0.0;
    {
        double spl[W];
        for (int l = 0; l < W; ++l) {
            double s = 0.0;
            for (int j = 0; j < J; ++j) s += Spoly[j][l];
            spl[l] = s;
        }
        const double SP = warp_sum(spl);
        double sumF2 = 0.0;
        /* F2's gradient terms are collected in their own accumulator and added to the cost
           gradient once, as the kernel does (it meets the hinges while it is still summing
           the ellipse cost terms) */
        lanes_t fx, fy;
        for (int j = 0; j < J; ++j)
            for (int l = 0; l < W; ++l) { fx[j][l] = 0.0; fy[j][l] = 0.0; }
        if (S->Ndyn == 0) {
            f2sq = SP * SP;
            sumF2 = SP;
            if (F2out) F2out[0] = SP;
        }
        for (int i = 0; i < S->Ndyn; ++i) {
            double hl[W];
            static __thread ell_t A[JMAX][W], B[JMAX][W];
            int any = 0;
            for (int l = 0; l < W; ++l) {
                double h = 0.0;
                for (int j = 0; j < J; ++j) {
                    const int k = act[j][l] ? kk[j][l] : N - 1;
                    A[j][l].hr = 0.0; B[j][l].hr = 0.0;
                    if (act[j][l]) {
                        ellipse_terms(GRAD, S->e0 + i, S->Ndyn, X[j][l], Y[j][l], &A[j][l]);
                        ellipse_terms(GRAD, S->et + k + i * N, S->Ndyn * N, X[j][l], Y[j][l], &B[j][l]);
                    }
                    h += A[j][l].hr + B[j][l].hr;
                }
                hl[l] = h;
                any |= h > 0.0;
            }
            double F2i = SP;
            if (any) F2i += warp_sum(hl);
            if (F2out) F2out[i] = F2i;
            f2sq = fma(F2i, F2i, f2sq);
            sumF2 += F2i;
            if (GRAD && F2i > 0.0 && S->G == 0) {
                const double m = c * F2i;
                for (int j = 0; j < J; ++j)
                    for (int l = 0; l < W; ++l) {
                        if (A[j][l].hr > 0.0) { fx[j][l] = fma(m, A[j][l].hrx, fx[j][l]); fy[j][l] = fma(m, A[j][l].hry, fy[j][l]); }
                        if (B[j][l].hr > 0.0) { fx[j][l] = fma(m, B[j][l].hrx, fx[j][l]); fy[j][l] = fma(m, B[j][l].hry, fy[j][l]); }
                    }
            }
            if (GRAD && F2i > 0.0 && S->G > 0 && any) {
                /* team mode: an ellipse with a raw hinge somewhere contributes at every step, both
                   slots added first (hrx, hry are zero where there is no hinge) */
                const double m = c * F2i;
                for (int j = 0; j < J; ++j)
                    for (int l = 0; l < W; ++l)
                        if (act[j][l]) {
                            fx[j][l] = fma(m, A[j][l].hrx + B[j][l].hrx, fx[j][l]);
                            fy[j][l] = fma(m, A[j][l].hry + B[j][l].hry, fy[j][l]);
                        }
            }
        }
        if (GRAD) {
            const double m = c * sumF2;
            for (int j = 0; j < J; ++j)
                for (int l = 0; l < W; ++l) {
                    gx[j][l] = fma(m, dSx[j][l], gx[j][l] + fx[j][l]);
                    gy[j][l] = fma(m, dSy[j][l], gy[j][l] + fy[j][l]);
                }
        }
    }
    /* accelerations */
    lanes_t gFa, gFw;
    double dist2[W];
    for (int l = 0; l < W; ++l) dist2[l] = 0.0;
    {
        const double accp = q[8], waccp = q[9];
        const double cdiv = fmax(c, 1.0);
        for (int j = 0; j < J; ++j)
            for (int l = 0; l < W; ++l) {
                double vp, wp;
                if (l > 0) { vp = v[j][l - 1]; wp = w[j][l - 1]; }
                else if (j == 0) { vp = S->um1[0]; wp = S->um1[1]; }
                else { vp = v[j - 1][W - 1]; wp = w[j - 1][W - 1]; }
                const double acc = (v[j][l] - vp) * S->inv_ts, wacc = (w[j][l] - wp) * S->inv_ts;
                const double za = acc + ya[j][l] / cdiv, zw = wacc + yw[j][l] / cdiv;
                const double ra = za > S->amax ? za - S->amax : (za < S->amin ? za - S->amin : 0.0);
                const double rwv = zw > S->wamax ? zw - S->wamax : (zw < -S->wamax ? zw + S->wamax : 0.0);
                if (act[j][l]) {
                    cost[l] = fma(accp, acc * acc, cost[l]);
                    cost[l] = fma(waccp, wacc * wacc, cost[l]);
                    dist2[l] = fma(ra, ra, fma(rwv, rwv, dist2[l]));
                    gFa[j][l] = fma(2.0 * accp, acc, c * ra);
                    gFw[j][l] = fma(2.0 * waccp, wacc, c * rwv);
                    if (F1out) { F1out[kk[j][l]] = acc; F1out[N + kk[j][l]] = wacc; }
                } else {
                    gFa[j][l] = 0.0; gFw[j][l] = 0.0;
                }
            }
    }
    /* one reduction for psi, as the kernel does: each lane folds its share of the ALM distance
       term into its stage cost before the butterfly; f is reduced separately (for c = 0 the
       two sums have the same bits) */
    const double hc = 0.5 * c;
    double pl[W];
    for (int l = 0; l < W; ++l) pl[l] = fma(hc, dist2[l], cost[l]);
    const double ps = warp_sum(pl);
    out->f = warp_sum(cost);
    out->f2sq = f2sq;
    out->psi = fma(hc, f2sq, ps);
    if (GRAD) {
        lanes_t Gx, Gy, hh, Hex;
        const double gthN_all = warp_sum(gthN);
        suffix(J, gx, 0, Gx);
        suffix(J, gy, 0, Gy);
        for (int j = 0; j < J; ++j)
            for (int l = 0; l < W; ++l) hh[j][l] = fma(Gy[j][l], dx[j][l], -(Gx[j][l] * dy[j][l]));
        suffix(J, hh, Hex, 0);
        for (int j = J - 1; j >= 0; --j)
            for (int l = 0; l < W; ++l) {
                double na, nw;
                if (l < W - 1) { na = gFa[j][l + 1]; nw = gFw[j][l + 1]; }
                else if (j == J - 1) { na = 0.0; nw = 0.0; }
                else { na = gFa[j + 1][0]; nw = gFw[j + 1][0]; }
                const double kv = S->k6 * v[j][l] * S->ts;
                const double dxw = -kv * fma(2.0, sb[j][l], sc[j][l]);
                const double dyw = kv * fma(2.0, cb[j][l], cc[j][l]);
                double g0 = gvd[j][l] + S->k6 * fma(Gx[j][l], Cc[j][l], Gy[j][l] * Ss[j][l]) + (gFa[j][l] - na) * S->inv_ts;
                double g1 = gwd[j][l] + fma(Gx[j][l], dxw, Gy[j][l] * dyw) + S->ts * (Hex[j][l] + gthN_all) +
                            (gFw[j][l] - nw) * S->inv_ts;
                out->gv[j][l] = act[j][l] ? g0 : 0.0;
                out->gw[j][l] = act[j][l] ? g1 : 0.0;
            }
    }
}

static void to_lanes(int N, int J, const double* u, lanes_t v, lanes_t w)
{
    for (int j = 0; j < J; ++j)
        for (int l = 0; l < W; ++l) {
            const int k = l + W * j;
            v[j][l] = (u && k < N) ? u[2 * k] : 0.0;
            w[j][l] = (u && k < N) ? u[2 * k + 1] : 0.0;
        }
}
static void to_lanes_y(int N, int J, const double* y, lanes_t a, lanes_t b)
{
    for (int j = 0; j < J; ++j)
        for (int l = 0; l < W; ++l) {
            const int k = l + W * j;
            a[j][l] = (y && k < N) ? y[k] : 0.0;
            b[j][l] = (y && k < N) ? y[N + k] : 0.0;
        }
}

int32_t mpcl_eval(const mpcb_dims* d, const mpcb_robot* rb, const double* p, const double* u,
                  const double* y, double c, double* f_out, double* psi_out, double* grad,
                  double* F1_out, double* F2_out)
{
    scen_t S;
    int rc = scen_stage(&S, d, rb, p);
    if (rc) return rc;
    lanes_t v, w, ya, yw;
    to_lanes(S.N, S.J, u, v, w);
    to_lanes_y(S.N, S.J, y, ya, yw);
    eval_out_t o;
    eval_psi(&S, v, w, c, ya, yw, 1, &o, F1_out, F2_out);
    if (f_out) *f_out = o.f;
    if (psi_out) *psi_out = o.psi;
    if (grad)
        for (int k = 0; k < S.N; ++k) {
            grad[2 * k] = o.gv[k / W][k % W];
            grad[2 * k + 1] = o.gw[k / W][k % W];
        }
    scen_free(&S);
    return MPCB_OK;
}

/* ================================ solver (kernel order) ===================== */
typedef struct {
    int N, J, mem, M, head, active, first_old;
    double gamma;
    double s[MPCB_MAX_LBFGS + 1][2 * MAXN], y[MPCB_MAX_LBFGS + 1][2 * MAXN];
    double rho[MPCB_MAX_LBFGS + 1], alpha[MPCB_MAX_LBFGS + 1];
    lanes_t os0, os1, og0, og1;
} lb_t;

typedef struct {
    lanes_t u0, u1, g0, g1, gp0, gp1, h0, h1, r0, r1, d0, d1, s0, s1, ya, yw;
    double c, gamma, sigma, Lc, cost, norm_r, akkt_tol;
    int iter, n_cost, n_grad, n_small;
} inst_t;

#define FORJL for (int j = 0; j < J; ++j) for (int l = 0; l < W; ++l)

static void grad_and_half_step(const scen_t* S, inst_t* I, lanes_t p0, lanes_t p1)
{
    const int J = S->J;
    FORJL {
        I->s0[j][l] = fma(-I->gamma, I->g0[j][l], p0[j][l]);
        I->s1[j][l] = fma(-I->gamma, I->g1[j][l], p1[j][l]);
        const double a0 = I->s0[j][l], a1 = I->s1[j][l];
        I->h0[j][l] = a0 < S->vmin ? S->vmin : (a0 > S->vmax ? S->vmax : a0);
        I->h1[j][l] = a1 < -S->wmax ? -S->wmax : (a1 > S->wmax ? S->wmax : a1);
    }
}
static void compute_fpr(const scen_t* S, inst_t* I)
{
    const int J = S->J;
    FORJL { I->r0[j][l] = I->u0[j][l] - I->h0[j][l]; I->r1[j][l] = I->u1[j][l] - I->h1[j][l]; }
    I->norm_r = sqrt(sumsq2(J, I->r0, I->r1));
}
static int phys(const lb_t* B, int logical)
{
    int r = B->head + logical;
    return r >= B->M ? r - B->M : r;
}
static void lbfgs_update(const mpcb_solver_cfg* cfg, const scen_t* S, lb_t* B, const inst_t* I)
{
    const int J = S->J, N = S->N;
    if (B->first_old) {
        B->first_old = 0;
        FORJL { B->os0[j][l] = I->u0[j][l]; B->os1[j][l] = I->u1[j][l]; B->og0[j][l] = I->r0[j][l]; B->og1[j][l] = I->r1[j][l]; }
        return;
    }
    lanes_t sn0, sn1, yn0, yn1;
    double pys[W], pss[W], pyy[W];
    for (int l = 0; l < W; ++l) {
        double ys = 0.0, ss = 0.0, yy = 0.0;
        for (int j = 0; j < J; ++j) {
            sn0[j][l] = I->u0[j][l] - B->os0[j][l]; sn1[j][l] = I->u1[j][l] - B->os1[j][l];
            yn0[j][l] = I->r0[j][l] - B->og0[j][l]; yn1[j][l] = I->r1[j][l] - B->og1[j][l];
            ys = fma(sn0[j][l], yn0[j][l], fma(sn1[j][l], yn1[j][l], ys));
            ss = fma(sn0[j][l], sn0[j][l], fma(sn1[j][l], sn1[j][l], ss));
            yy = fma(yn0[j][l], yn0[j][l], fma(yn1[j][l], yn1[j][l], yy));
        }
        pys[l] = ys; pss[l] = ss; pyy[l] = yy;
    }
    const double ys = warp_sum(pys), ss = warp_sum(pss), yy = warp_sum(pyy);
    int ok;
    if (ss <= 2.2250738585072014e-308 || (cfg->sy_epsilon > 0.0 && ys <= cfg->sy_epsilon)) {
        ok = 0;
    } else if (cfg->cbfgs_epsilon > 0.0 && cfg->cbfgs_alpha > 0.0) {
        const double lhs = ys / ss;
        const double rhs = cfg->cbfgs_epsilon * (cfg->cbfgs_alpha == 1.0 ? I->norm_r : pow(I->norm_r, cfg->cbfgs_alpha));
        ok = lhs > rhs && isfinite(lhs) && isfinite(rhs);
    } else {
        ok = 1;
    }
    if (!ok) return;
    FORJL { B->os0[j][l] = I->u0[j][l]; B->os1[j][l] = I->u1[j][l]; B->og0[j][l] = I->r0[j][l]; B->og1[j][l] = I->r1[j][l]; }
    B->head = B->head == 0 ? B->M - 1 : B->head - 1;
    FORJL {
        const int k = l + W * j;
        if (k < N) {
            B->s[B->head][k] = sn0[j][l]; B->s[B->head][N + k] = sn1[j][l];
            B->y[B->head][k] = yn0[j][l]; B->y[B->head][N + k] = yn1[j][l];
        }
    }
    B->rho[B->head] = 1.0 / ys;
    B->gamma = ys / yy;
    B->active = B->active + 1 < B->mem ? B->active + 1 : B->mem;
}
static void lbfgs_apply(const scen_t* S, lb_t* B, lanes_t q0, lanes_t q1)
{
    const int J = S->J, N = S->N;
    if (B->active == 0) return;
    for (int kq = 0; kq < B->active; ++kq) {
        const int row = phys(B, kq);
        double part[W];
        for (int l = 0; l < W; ++l) {
            double s = 0.0;
            for (int j = 0; j < J; ++j) {
                const int k = l + W * j;
                const double sv0 = k < N ? B->s[row][k] : 0.0, sv1 = k < N ? B->s[row][N + k] : 0.0;
                s = fma(sv0, q0[j][l], fma(sv1, q1[j][l], s));
            }
            part[l] = s;
        }
        const double a = B->rho[row] * warp_sum(part);
        B->alpha[row] = a;
        FORJL {
            const int k = l + W * j;
            const double yv0 = k < N ? B->y[row][k] : 0.0, yv1 = k < N ? B->y[row][N + k] : 0.0;
            q0[j][l] = fma(-a, yv0, q0[j][l]);
            q1[j][l] = fma(-a, yv1, q1[j][l]);
        }
    }
    FORJL { q0[j][l] *= B->gamma; q1[j][l] *= B->gamma; }
    for (int kq = B->active - 1; kq >= 0; --kq) {
        const int row = phys(B, kq);
        double part[W];
        for (int l = 0; l < W; ++l) {
            double s = 0.0;
            for (int j = 0; j < J; ++j) {
                const int k = l + W * j;
                const double y0 = k < N ? B->y[row][k] : 0.0, y1 = k < N ? B->y[row][N + k] : 0.0;
                s = fma(y0, q0[j][l], fma(y1, q1[j][l], s));
            }
            part[l] = s;
        }
        const double beta = B->rho[row] * warp_sum(part);
        const double cf = B->alpha[row] - beta;
        FORJL {
            const int k = l + W * j;
            const double sv0 = k < N ? B->s[row][k] : 0.0, sv1 = k < N ? B->s[row][N + k] : 0.0;
            q0[j][l] = fma(cf, sv0, q0[j][l]);
            q1[j][l] = fma(cf, sv1, q1[j][l]);
        }
    }
}

static void eval_at(const scen_t* S, inst_t* I, lanes_t p0, lanes_t p1, double ceff, int grad,
                    eval_out_t* o)
{
    eval_psi(S, p0, p1, ceff, I->ya, I->yw, grad, o, 0, 0);
    if (grad) I->n_grad++; else I->n_cost++;
}

/* PANOCEngine::step in kernel order; returns 1 to continue */
static int panoc_step(const mpcb_solver_cfg* cfg, const scen_t* S, inst_t* I, lb_t* B)
{
    const int J = S->J;
    eval_out_t o;
    if (I->iter >= 1) FORJL { I->gp0[j][l] = I->g0[j][l]; I->gp1[j][l] = I->g1[j][l]; }
    compute_fpr(S, I);
    if (I->norm_r < cfg->tolerance) {
        double p[W];
        for (int l = 0; l < W; ++l) {
            double a = 0.0;
            for (int j = 0; j < J; ++j) {
                const double t0 = I->r0[j][l] / I->gamma + I->g0[j][l] - I->gp0[j][l];
                const double t1 = I->r1[j][l] / I->gamma + I->g1[j][l] - I->gp1[j][l];
                a = fma(t0, t0, fma(t1, t1, a));
            }
            p[l] = a;
        }
        if (sqrt(warp_sum(p)) < I->akkt_tol) return 0;
        I->n_small++;   /* |gamma fpr| < tolerance, AKKT test failed: the solve goes on */
    }
    eval_at(S, I, I->h0, I->h1, I->c, 0, &o);
    double cost_half = o.psi;
    int it_lip = 0;
    if (I->iter == 0) {
        eval_at(S, I, I->u0, I->u1, I->c, 0, &o);
        I->cost = o.psi;
    }
    for (;;) {
        const double ip = dotw(J, I->g0, I->g1, I->r0, I->r1);
        const double rhs = I->cost + 1e-6 * fabs(I->cost) - ip + (0.95 / (2.0 * I->gamma)) * (I->norm_r * I->norm_r);
        if (!(cost_half > rhs && it_lip < 10 && I->Lc < 1e9)) break;
        B->active = 0; B->first_old = 1;
        I->Lc *= 2.0;
        I->gamma /= 2.0;
        grad_and_half_step(S, I, I->u0, I->u1);
        eval_at(S, I, I->h0, I->h1, I->c, 0, &o);
        cost_half = o.psi;
        compute_fpr(S, I);
        ++it_lip;
    }
    I->sigma = (1.0 - 0.95) / (4.0 * I->gamma);
    lbfgs_update(cfg, S, B, I);
    if (I->iter > 0) {
        FORJL { I->d0[j][l] = I->r0[j][l]; I->d1[j][l] = I->r1[j][l]; }
        lbfgs_apply(S, B, I->d0, I->d1);
    }
    if (I->iter == 0) {
        FORJL { I->u0[j][l] = I->h0[j][l]; I->u1[j][l] = I->h1[j][l]; }
        eval_at(S, I, I->u0, I->u1, I->c, 1, &o);
        I->cost = o.psi;
        FORJL { I->g0[j][l] = o.gv[j][l]; I->g1[j][l] = o.gw[j][l]; }
        grad_and_half_step(S, I, I->u0, I->u1);
    } else {
        double pd[W];
        for (int l = 0; l < W; ++l) {
            double dd = 0.0;
            for (int j = 0; j < J; ++j) {
                const double e0 = I->s0[j][l] - I->h0[j][l], e1 = I->s1[j][l] - I->h1[j][l];
                dd = fma(e0, e0, fma(e1, e1, dd));
            }
            pd[l] = dd;
        }
        const double dist2 = warp_sum(pd);
        const double fbe = I->cost - 0.5 * I->gamma * sumsq2(J, I->g0, I->g1) + 0.5 * dist2 / I->gamma;
        const double rhs_ls = fbe - I->sigma * (I->norm_r * I->norm_r);
        double tau = 1.0;
        int ls = 0;
        lanes_t p0, p1;
        for (;;) {
            FORJL {
                p0[j][l] = I->u0[j][l] - (1.0 - tau) * I->r0[j][l] - tau * I->d0[j][l];
                p1[j][l] = I->u1[j][l] - (1.0 - tau) * I->r1[j][l] - tau * I->d1[j][l];
            }
            eval_at(S, I, p0, p1, I->c, 1, &o);
            I->cost = o.psi;
            FORJL { I->g0[j][l] = o.gv[j][l]; I->g1[j][l] = o.gw[j][l]; }
            grad_and_half_step(S, I, p0, p1);
            for (int l = 0; l < W; ++l) {
                double dd = 0.0;
                for (int j = 0; j < J; ++j) {
                    const double e0 = I->s0[j][l] - I->h0[j][l], e1 = I->s1[j][l] - I->h1[j][l];
                    dd = fma(e0, e0, fma(e1, e1, dd));
                }
                pd[l] = dd;
            }
            const double d2 = warp_sum(pd);
            const double lhs = I->cost - 0.5 * I->gamma * sumsq2(J, I->g0, I->g1) + 0.5 * d2 / I->gamma;
            if (!(lhs > rhs_ls && ls < 10)) break;
            tau /= 2.0;
            ++ls;
        }
        FORJL { I->u0[j][l] = p0[j][l]; I->u1[j][l] = p1[j][l]; }
    }
    I->iter++;
    return 1;
}

/* out_scalars as mpco_solve: {cost, fpr, f1_infeas, f2_norm, penalty, n_outer, n_inner,
 * n_cost, n_grad, exit_status, n_small}; n_cost / n_grad count the kernel's evaluations (cost-only,
 * cost+gradient), n_small the inner iterations that met |gamma fpr| < tolerance but not the AKKT
 * test. */
int32_t mpcl_solve(const mpcb_dims* d, const mpcb_robot* rb, const mpcb_solver_cfg* cfg,
                   const double* p, const double* u0, const double* y0, const double* c0,
                   double* u_out, double* y_out, double* out_scalars)
{
    if (!d || !rb || !cfg || !p || !u_out) return MPCB_E_NULL;
    scen_t Sc;
    g_force_team = cfg->team_mode == 1;
    int rc = scen_stage(&Sc, d, rb, p);
    g_force_team = 0;
    if (rc) return rc;
    const scen_t* S = &Sc;
    const int N = S->N, J = S->J;
    inst_t* I = (inst_t*)calloc(1, sizeof(inst_t));
    lb_t* B = (lb_t*)calloc(1, sizeof(lb_t));
    B->N = N; B->J = J; B->mem = cfg->lbfgs_mem; B->M = cfg->lbfgs_mem + 1;
    B->head = 0; B->active = 0; B->first_old = 1; B->gamma = 1.0;
    to_lanes(N, J, u0, I->u0, I->u1);
    to_lanes_y(N, J, y0, I->ya, I->yw);
    I->c = c0 ? *c0 : cfg->initial_penalty;
    I->akkt_tol = cfg->initial_tolerance;
    lanes_t yp_a, yp_w;
    memset(yp_a, 0, sizeof(yp_a)); memset(yp_w, 0, sizeof(yp_w));
    double dy = 0.0, dy_plus = 0.0, f2n = 0.0, f2n_plus = 0.0, last_fpr = -1.0, fcost = 0.0;
    int alm_iter = 0, n_outer = 0, inner_total = 0, status = MPCB_CONVERGED, failed = 0;
    const double EPS = 2.220446049250313e-16;
    eval_out_t o;

    const int budget = cfg->max_inner_total > 0 ? cfg->max_inner_total : 0;
    for (int outer = 1; outer <= cfg->max_outer; ++outer) {
        /* AlmOptimizer::solve: no time left -> NotConvergedOutOfTime (the clock is the inner
           iteration count, cfg->max_inner_total) */
        if (budget > 0 && inner_total >= budget) { status = MPCB_NOT_CONVERGED_OUT_OF_TIME; break; }
        ++n_outer;
        FORJL {
            I->ya[j][l] = fmin(fmax(I->ya[j][l], -1e12), 1e12);
            I->yw[j][l] = fmin(fmax(I->yw[j][l], -1e12), 1e12);
            I->gp0[j][l] = 0.0; I->gp1[j][l] = 0.0;
        }
        /* PANOCEngine::init */
        B->active = 0; B->first_old = 1;
        I->iter = 0;
        eval_at(S, I, I->u0, I->u1, I->c, 1, &o);
        I->cost = o.psi;
        {
            double ph[W];
            for (int l = 0; l < W; ++l) {
                double hs = 0.0;
                for (int j = 0; j < J; ++j) {
                    const int act = l + W * j < N;
                    I->g0[j][l] = o.gv[j][l]; I->g1[j][l] = o.gw[j][l];
                    const double h0 = act ? ((1e-6 * I->u0[j][l] > 1e-12) ? 1e-6 * I->u0[j][l] : 1e-12) : 0.0;
                    const double h1 = act ? ((1e-6 * I->u1[j][l] > 1e-12) ? 1e-6 * I->u1[j][l] : 1e-12) : 0.0;
                    hs = fma(h0, h0, fma(h1, h1, hs));
                    I->u0[j][l] += h0; I->u1[j][l] += h1;
                }
                ph[l] = hs;
            }
            const double norm_h = sqrt(warp_sum(ph));
            eval_at(S, I, I->u0, I->u1, I->c, 1, &o);
            lanes_t t0, t1;
            FORJL { t0[j][l] = o.gv[j][l] - I->g0[j][l]; t1[j][l] = o.gw[j][l] - I->g1[j][l]; }
            I->Lc = sqrt(sumsq2(J, t0, t1)) / norm_h;
        }
        I->gamma = 0.95 / fmax(I->Lc, 1e-10);
        I->sigma = (1.0 - 0.95) / (4.0 * I->gamma);
        grad_and_half_step(S, I, I->u0, I->u1);
        /* PANOCOptimizer::solve */
        int num_iter = 0, cont = 1;
        int flag = panoc_step(cfg, S, I, B);
        while (flag && cont) {
            ++num_iter;
            cont = num_iter < cfg->max_inner && (budget <= 0 || inner_total + num_iter < budget);
            flag = panoc_step(cfg, S, I, B);
        }
        int fin = 1;
        FORJL fin = fin && isfinite(I->u0[j][l]) && isfinite(I->u1[j][l]);
        if (!fin) { status = MPCB_NOT_FINITE_COMPUTATION; failed = 1; break; }
        FORJL { I->u0[j][l] = I->h0[j][l]; I->u1[j][l] = I->h1[j][l]; }
        const int inner = cont ? MPCB_CONVERGED
                               : (num_iter >= cfg->max_inner ? MPCB_NOT_CONVERGED_ITERATIONS
                                                             : MPCB_NOT_CONVERGED_OUT_OF_TIME);
        last_fpr = I->norm_r;
        inner_total += num_iter;
        eval_at(S, I, I->u0, I->u1, 0.0, 0, &o);
        fcost = o.f;
        f2n_plus = sqrt(o.f2sq);
        {
            double pd[W];
            for (int l = 0; l < W; ++l) {
                double dsum = 0.0;
                for (int j = 0; j < J; ++j) {
                    const int act = l + W * j < N;
                    double vp, wp;
                    if (l > 0) { vp = I->u0[j][l - 1]; wp = I->u1[j][l - 1]; }
                    else if (j == 0) { vp = S->um1[0]; wp = S->um1[1]; }
                    else { vp = I->u0[j - 1][W - 1]; wp = I->u1[j - 1][W - 1]; }
                    const double acc = (I->u0[j][l] - vp) * S->inv_ts, wacc = (I->u1[j][l] - wp) * S->inv_ts;
                    const double za = acc + I->ya[j][l] / I->c, zw = wacc + I->yw[j][l] / I->c;
                    const double pa = fmin(fmax(za, S->amin), S->amax), pw = fmin(fmax(zw, -S->wamax), S->wamax);
                    yp_a[j][l] = act ? I->ya[j][l] + I->c * (acc - pa) : 0.0;
                    yp_w[j][l] = act ? I->yw[j][l] + I->c * (wacc - pw) : 0.0;
                    const double e0 = yp_a[j][l] - I->ya[j][l], e1 = yp_w[j][l] - I->yw[j][l];
                    dsum = fma(e0, e0, fma(e1, e1, dsum));
                }
                pd[l] = dsum;
            }
            dy_plus = sqrt(warp_sum(pd));
        }
        const int c1 = alm_iter > 0 && dy_plus <= I->c * cfg->delta_tolerance + EPS;
        const int c2 = f2n_plus <= cfg->delta_tolerance + EPS;
        const int c3 = I->akkt_tol <= cfg->tolerance + EPS;
        if (c1 && c2 && c3) { status = inner; break; }
        const int stall = alm_iter == 0 || (dy_plus <= cfg->sufficient_decrease * dy + EPS &&
                                            f2n_plus <= cfg->sufficient_decrease * f2n + EPS);
        if (!stall) I->c *= cfg->penalty_update;
        I->akkt_tol = fmax(I->akkt_tol * cfg->inner_tol_update, cfg->tolerance);
        ++alm_iter;
        dy = dy_plus;
        f2n = f2n_plus;
        FORJL { I->ya[j][l] = yp_a[j][l]; I->yw[j][l] = yp_w[j][l]; }
    }
    if (!failed && n_outer == cfg->max_outer) status = MPCB_NOT_CONVERGED_ITERATIONS;
    for (int k = 0; k < N; ++k) {
        u_out[2 * k] = I->u0[k / W][k % W];
        u_out[2 * k + 1] = I->u1[k / W][k % W];
        if (y_out) { y_out[k] = yp_a[k / W][k % W]; y_out[N + k] = yp_w[k / W][k % W]; }
    }
    if (out_scalars) {
        out_scalars[0] = failed ? NAN : fcost;
        out_scalars[1] = last_fpr;
        out_scalars[2] = dy_plus / I->c;
        out_scalars[3] = f2n_plus;
        out_scalars[4] = I->c;
        out_scalars[5] = n_outer;
        out_scalars[6] = inner_total;
        out_scalars[7] = I->n_cost;
        out_scalars[8] = I->n_grad;
        out_scalars[9] = status;
        out_scalars[10] = I->n_small;
    }
    free(I); free(B);
    scen_free(&Sc);
    return MPCB_OK;
}

int32_t mpcl_solve_batch(const mpcb_dims* d, const mpcb_robot* rb, const mpcb_solver_cfg* cfg,
                         int32_t n_p, int32_t starts, const double* p, const double* u0,
                         double* u_out, double* out_scalars, int32_t np, int32_t threads)
{
    const int n = 2 * d->N;
    const long B = (long)n_p * starts;
    int rc = 0;
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads > 0 ? threads : 1)
#endif
    for (long b = 0; b < B; ++b) {
        int r = mpcl_solve(d, rb, cfg, p + (b / starts) * np, u0 ? u0 + b * n : 0, 0, 0, u_out + b * n,
                           0, out_scalars ? out_scalars + b * 11 : 0);
        if (r) rc = r;
    }
    return rc;
}
