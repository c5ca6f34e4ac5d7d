/*
 * mpc_oracle.c — CPU restatement (plain C, IEEE f64, one instance at a time)
 * of the reference's NMPC hot path.  TEST INFRASTRUCTURE ONLY: nothing in the
 * product (dyobav_mpcnwta_warehouse_b200/) may import, link or call this; only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs do.
 *
 * Two halves, pinned differently:
 *
 *  (1) Problem definition — f(u;p), grad f, F1, F2 and the augmented cost psi.
 *      Follows /root/reference/src/pkg_mpc_tracker/solver_build/
 *      mpc_builder.py:45-174, mpc_cost.py:6-95, mpc_helper.py:5-75 and
 *      src/basic_motion_model/motion_model.py:141-163 term by term, in the
 *      reference's own operation order.  PINNED: tests/golden/ (npz files) are
 *      produced by executing those reference files themselves (numeric
 *      stand-in for casadi, see tests/golden/gen_golden.py) and this file must
 *      reproduce them; plus the reference's own known-answer values
 *      (src/tests/test_mpc_builder.py:16-253).
 *
 *  (2) The solver — OpEn's ALM/PM outer loop around PANOC with L-BFGS.  That
 *      code is a third-party dependency absent from /root/reference
 *      (opengen==0.6.13 -> Rust crate optimization_engine 0.7.x + lbfgs 0.2.x,
 *      requirements.txt:7) and cannot be built here (no cargo, no casadi).
 *      It is restated from the published algorithm / upstream sources from
 *      memory: PARITY UNPINNED for this half.  Each routine names the upstream
 *      function it restates.
 *
 * Build: make -C oracle   (gcc -O2 -ffp-contract=off: no FMA contraction so the
 * arithmetic is the plain IEEE sequence the comments describe).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../include/mpcb.h"

#define NU 2
#define NS 3
#define NQ 10
#define NDYNPAR 6

typedef struct {
    int u_m1, s_0, s_N, q, r_s, r_v, c_0, c, o_s, o_d, q_stc, q_dyn, np;
} layout_t;

/* mpc_builder.py:47-60: z = [u_m1,s_0,s_N,q,r_s,r_v,c_0,c,o_s,o_d,q_stc,q_dyn] */
static layout_t make_layout(const mpcb_dims* d)
{
    layout_t L;
    int N = d->N;
    L.u_m1 = 0;
    L.s_0 = L.u_m1 + NU;
    L.s_N = L.s_0 + NS;
    L.q = L.s_N + NS;
    L.r_s = L.q + NQ;
    L.r_v = L.r_s + NS * N;
    L.c_0 = L.r_v + N;
    L.c = L.c_0 + NS * d->Nother;
    L.o_s = L.c + NS * N * d->Nother;
    L.o_d = L.o_s + 3 * d->nedge * d->Nstc;
    L.q_stc = L.o_d + NDYNPAR * (N + 1) * d->Ndyn;
    L.q_dyn = L.q_stc + N;
    L.np = L.q_dyn + N;
    return L;
}

int32_t mpco_param_len(const mpcb_dims* d) { return make_layout(d).np; }
static int n2_of(const mpcb_dims* d) { return d->Ndyn > 0 ? d->Ndyn : 1; }

/* CasADi's derivative convention for fmax(0,x) wrt x: 1 / 0.5 (tie) / 0. */
static double dmax0(double x) { return x > 0.0 ? 1.0 : (x == 0.0 ? 0.5 : 0.0); }

/* ---- mpc_helper.py:38-52 inside_ellipses, one ellipse; cos/sin of the angle
 *      are passed in so the raw and inflated evaluations share them --------- */
static double ellipse_ind_cs(double x, double y, double cx, double cy, double rx,
                             double ry, double ca, double sa, double* dEdx, double* dEdy)
{
    double a = (x - cx) * ca + (y - cy) * sa;
    double b = (x - cx) * sa - (y - cy) * ca;
    double R1 = (rx + 1e-6) * (rx + 1e-6);
    double R2 = (ry + 1e-6) * (ry + 1e-6);
    double E = 1.0 - (a * a) / R1 - (b * b) / R2;
    if (dEdx) {
        *dEdx = -(2.0 * a * ca) / R1 - (2.0 * b * sa) / R2;
        *dEdy = -(2.0 * a * sa) / R1 + (2.0 * b * ca) / R2;
    }
    return E;
}
static double ellipse_ind(double x, double y, double cx, double cy, double rx,
                          double ry, double ang, double* dEdx, double* dEdy)
{
    return ellipse_ind_cs(x, y, cx, cy, rx, ry, cos(ang), sin(ang), dEdx, dEdy);
}

/* ---- mpc_helper.py:54-75 inside_cvx_polygon ------------------------------- */
static double polygon_ind(double x, double y, const double* b, const double* a0,
                          const double* a1, int nedge, double* dIdx, double* dIdy)
{
    double m[MPCB_MAX_EDGE], r[MPCB_MAX_EDGE];
    double I = 1.0;
    for (int e = 0; e < nedge; ++e) {
        r[e] = b[e] + (-a0[e]) * x + (-a1[e]) * y; /* mtimes([b,-a0,-a1]^T,[1,x,y]) */
        m[e] = fmax(0.0, r[e]);
        I *= m[e];
    }
    if (dIdx) {
        double gx = 0.0, gy = 0.0;
        for (int e = 0; e < nedge; ++e) {
            double prod = 1.0;
            for (int j = 0; j < nedge; ++j)
                if (j != e) prod *= m[j];
            double w = prod * dmax0(r[e]);
            gx += w * (-a0[e]);
            gy += w * (-a1[e]);
        }
        *dIdx = gx;
        *dIdy = gy;
    }
    return I;
}

/* ---- mpc_helper.py:17-36 dist_to_lineseg, then **2 (mpc_cost.py:93) -------- */
static double seg_dist_sq(double px, double py, double s1x, double s1y, double s2x,
                          double s2y, double* gx, double* gy)
{
    double dx = s2x - s1x, dy = s2y - s1y;
    double den = dx * dx + dy * dy + 1e-16;
    double t_hat = ((px - s1x) * dx + (py - s1y) * dy) / den;
    double t_in = fmax(t_hat, 0.0);
    double t_star = fmin(t_in, 1.0);
    double vx = s1x + t_star * dx - px;
    double vy = s1y + t_star * dy - py;
    double dist = sqrt(vx * vx + vy * vy);
    double d2 = dist * dist;
    if (gx) {
        /* d t_star / d t_hat with CasADi tie rules: fmax(t,0) then fmin(.,1) */
        double dmx = t_hat > 0.0 ? 1.0 : (t_hat == 0.0 ? 0.5 : 0.0);
        double dmn = t_in < 1.0 ? 1.0 : (t_in == 1.0 ? 0.5 : 0.0);
        double dt = dmx * dmn;
        double vd = vx * dx + vy * dy;
        /* d(d2)/dp = 2 v^T (d * dt_star/dp^T - I),  dt_hat/dp = d/den
         * (the reference differentiates sqrt then square: identical wherever
         * dist != 0; at dist == 0 CasADi yields NaN — SURVEY C-5 — the squared
         * form here yields 0, documented deviation) */
        *gx = 2.0 * (vd * dt * dx / den - vx);
        *gy = 2.0 * (vd * dt * dy / den - vy);
    }
    return d2;
}

/* ---- motion_model.py:141-163 unicycle_model, rk4=True ---------------------- */
static void rk4_step(const double s[3], double v, double w, double ts, double out[3])
{
    double k1[3], k2[3], k3[3], k4[3];
    k1[0] = ts * (v * cos(s[2]));
    k1[1] = ts * (v * sin(s[2]));
    k1[2] = ts * w;
    double th2 = s[2] + 0.5 * k1[2];
    k2[0] = ts * (v * cos(th2));
    k2[1] = ts * (v * sin(th2));
    k2[2] = ts * w;
    double th3 = s[2] + 0.5 * k2[2];
    k3[0] = ts * (v * cos(th3));
    k3[1] = ts * (v * sin(th3));
    k3[2] = ts * w;
    double th4 = s[2] + k3[2];
    k4[0] = ts * (v * cos(th4));
    k4[1] = ts * (v * sin(th4));
    k4[2] = ts * w;
    for (int i = 0; i < 3; ++i)
        out[i] = s[i] + (1.0 / 6.0) * (k1[i] + 2.0 * k2[i] + 2.0 * k3[i] + k4[i]);
}

/* Partials of the RK4 increment (dx,dy) wrt (v, w, theta); dtheta = ts*w. */
static void rk4_partials(double th, double v, double w, double ts, double* dx_dv,
                         double* dy_dv, double* dx_dw, double* dy_dw, double* dx_dth,
                         double* dy_dth)
{
    double tb = th + 0.5 * (ts * w), tc = th + ts * w;
    double C = cos(th) + 4.0 * cos(tb) + cos(tc);
    double S = sin(th) + 4.0 * sin(tb) + sin(tc);
    double k = ts / 6.0;
    *dx_dv = k * C;
    *dy_dv = k * S;
    *dx_dth = -k * v * S;
    *dy_dth = k * v * C;
    *dx_dw = -k * v * ts * (2.0 * sin(tb) + sin(tc));
    *dy_dw = k * v * ts * (2.0 * cos(tb) + cos(tc));
}

typedef struct {
    const mpcb_dims* d;
    const mpcb_robot* rb;
    const double* p;
    layout_t L;
} prob_t;

/*
 * Stage terms at state s_{k+1}=(x,y,th) reached with u_k=(v,w)
 * (mpc_builder.py:80-143).  mode 0: accumulate cost into *cost and penalty
 * pieces into S_poly / F2[] .  mode 1: given the F2 totals and penalty c,
 * return d(f + c/2 |F2|^2)/d(x,y) in g[0..1] and direct d/d(v,w) in gu[0..1].
 */
static void stage_terms(const prob_t* P, int k, double x, double y, double v, double w,
                        int mode, double* cost, double* F2, const double* F2tot,
                        double cpen, double g[2], double gu[2])
{
    const mpcb_dims* d = P->d;
    const layout_t* L = &P->L;
    const double* p = P->p;
    const int N = d->N;
    const double* q = p + L->q;
    const double qvel = q[1], rv = q[3], rw = q[4], qrpd = q[7];
    double gx = 0.0, gy = 0.0, gv = 0.0, gw = 0.0;
    double c_acc = 0.0;

    /* -- reference path deviation (mpc_cost.py:84-95): qrpd * mmin_i dist^2 to
     * segments i=k..N-1 of ref_states rows (row N duplicates row N-1,
     * mpc_builder.py:68-69).  mmin is CasADi's fold r=fmin(r,x_i) from +inf;
     * ties split the derivative 0.5/0.5 at each fold step. */
    {
        const double* rs = p + L->r_s;
        double wsel[MPCB_MAX_N], sgx[MPCB_MAX_N], sgy[MPCB_MAX_N];
        double r = INFINITY;
        int nseg = N - k;
        for (int j = 0; j < nseg; ++j) {
            int i = k + j;
            int i2 = (i + 1 < N) ? i + 1 : N - 1;
            double d2 = seg_dist_sq(x, y, rs[3 * i], rs[3 * i + 1], rs[3 * i2],
                                    rs[3 * i2 + 1], mode ? &sgx[j] : 0, mode ? &sgy[j] : 0);
            if (d2 < r) {
                for (int m = 0; m < j; ++m) wsel[m] = 0.0;
                wsel[j] = 1.0;
                r = d2;
            } else if (d2 == r) {
                for (int m = 0; m < j; ++m) wsel[m] *= 0.5;
                wsel[j] = 0.5;
            } else {
                wsel[j] = 0.0;
            }
        }
        c_acc += r * qrpd;
        if (mode)
            for (int j = 0; j < nseg; ++j)
                if (wsel[j] != 0.0) {
                    gx += qrpd * wsel[j] * sgx[j];
                    gy += qrpd * wsel[j] * sgy[j];
                }
    }
    /* -- reference speed (mpc_cost.py:78-79) and control (46-53) */
    {
        double rvk = p[L->r_v + k];
        c_acc += qvel * ((v - rvk) * (v - rvk));
        c_acc += rv * (v * v) + rw * (w * w);
        gv += 2.0 * qvel * (v - rvk) + 2.0 * rv * v;
        gw += 2.0 * rw * w;
    }
    /* -- fleet collision (mpc_cost.py:65-76): linear hinge on squared distance.
     * (i) c_0 robots 1..Nother-1 (slice starts at ns, mpc_builder.py:86-87), weight 1000
     * (ii) predicted c, robot-major c[r*3N + 3k + {0,1}] (:93-94), weight 10 */
    {
        double ds2 = P->rb->vehicle_width * P->rb->vehicle_width;
        double s1 = 0.0, s2 = 0.0;
        for (int r = 1; r < d->Nother; ++r) {
            double ox = p[L->c_0 + 3 * r], oy = p[L->c_0 + 3 * r + 1];
            double h = ds2 - ((x - ox) * (x - ox) + (y - oy) * (y - oy));
            s1 += fmax(0.0, h);
            if (mode) {
                double m = dmax0(h);
                gx += 1000.0 * m * (-2.0 * (x - ox));
                gy += 1000.0 * m * (-2.0 * (y - oy));
            }
        }
        for (int r = 0; r < d->Nother; ++r) {
            double ox = p[L->c + r * 3 * N + 3 * k], oy = p[L->c + r * 3 * N + 3 * k + 1];
            double h = ds2 - ((x - ox) * (x - ox) + (y - oy) * (y - oy));
            s2 += fmax(0.0, h);
            if (mode) {
                double m = dmax0(h);
                gx += 10.0 * m * (-2.0 * (x - ox));
                gy += 10.0 * m * (-2.0 * (y - oy));
            }
        }
        c_acc += 1000.0 * s1;
        c_acc += 10.0 * s2;
    }
    /* sum of F2 entries (the scalar polygon hinge is broadcast onto every one
     * of the Ndyn entries — SURVEY C-1) */
    double F2sum = 0.0;
    const int n2 = n2_of(d);
    if (mode)
        for (int i = 0; i < n2; ++i) F2sum += F2tot[i];

    /* -- static polygons (mpc_builder.py:100-108) */
    {
        double qs = p[L->q_stc + k];
        int ne = d->nedge;
        for (int i = 0; i < d->Nstc; ++i) {
            const double* e = p + L->o_s + i * 3 * ne;
            double dIx, dIy;
            double I = polygon_ind(x, y, e, e + ne, e + 2 * ne, ne, mode ? &dIx : 0,
                                   mode ? &dIy : 0);
            if (!mode) {
                double h = fmax(0.0, I);
                for (int j = 0; j < n2; ++j) F2[j] += h;
                c_acc += qs * (I * I);
            } else {
                double wI = 2.0 * qs * I + cpen * F2sum * dmax0(I);
                gx += wI * dIx;
                gy += wI * dIy;
            }
        }
    }
    /* -- dynamic ellipses: t=0 slot (mpc_builder.py:111-125) and t=k+1 slot
     * (:129-143).  o_d is obstacle-major, then time, then (x,y,rx,ry,ang,alpha). */
    {
        double vm = P->rb->vehicle_margin, sm = P->rb->social_margin;
        double qd = p[L->q_dyn + k];
        for (int pass = 0; pass < 2; ++pass) {
            int t = pass == 0 ? 0 : k + 1;
            double s_cost = 0.0;
            for (int i = 0; i < d->Ndyn; ++i) {
                const double* o = p + L->o_d + (i * (N + 1) + t) * NDYNPAR;
                double cx = o[0], cy = o[1], rx = o[2], ry = o[3], ang = o[4], al = o[5];
                double rxi = pass == 0 ? rx + vm + sm : rx + vm;
                double ryi = pass == 0 ? ry + vm + sm : ry + vm;
                double wgt = pass == 0 ? 1000.0 : qd;
                double dEx, dEy, dEix, dEiy;
                double ca = cos(ang), sa = sin(ang);
                double E = ellipse_ind_cs(x, y, cx, cy, rx, ry, ca, sa, mode ? &dEx : 0,
                                          mode ? &dEy : 0);
                double Ei = ellipse_ind_cs(x, y, cx, cy, rxi, ryi, ca, sa,
                                           mode ? &dEix : 0, mode ? &dEiy : 0);
                if (!mode) {
                    F2[i] += fmax(0.0, E);
                    double h = fmax(0.0, Ei);
                    s_cost += wgt * al * (h * h);
                } else {
                    double wE = cpen * F2tot[i] * dmax0(E);
                    double wEi = wgt * al * 2.0 * fmax(0.0, Ei) * dmax0(Ei);
                    gx += wE * dEx + wEi * dEix;
                    gy += wE * dEy + wEi * dEiy;
                }
            }
            c_acc += s_cost;
        }
    }
    if (!mode) {
        *cost += c_acc;
    } else {
        g[0] = gx;
        g[1] = gy;
        gu[0] = gv;
        gu[1] = gw;
    }
}

/* dist^2 to the rectangle C and the residual z - Proj_C(z) (opengen
 * Rectangle.distance_squared: fmax(0, fmax(z-zmax, zmin-z))^2 summed). */
static double rect_resid(double z, double lo, double hi)
{
    double a = z - hi, b = lo - z;
    double m = fmax(a, b);
    if (!(m > 0.0)) return 0.0;
    return a >= b ? m : -m; /* signed so that resid = z - proj(z) */
}

/*
 * Evaluate f, psi, grad psi, F1, F2 for one instance.
 *   y: n1 multipliers or NULL (0).  c: penalty.  Any output may be NULL.
 * psi = f + c/2*[dist2_C(F1 + y/max(c,1)) + |F2|^2]  (opengen builder,
 * "__construct_function_psi"; calling with c=0 returns f).
 */
int32_t mpco_eval(const mpcb_dims* d, const mpcb_robot* rb, const double* p,
                  const double* u, const double* y, double c, double* f_out,
                  double* psi_out, double* grad, double* F1_out, double* F2_out)
{
    if (!d || !rb || !p || !u) return MPCB_E_NULL;
    if (d->N < 1 || d->N > MPCB_MAX_N || d->nedge < 1 || d->nedge > MPCB_MAX_EDGE ||
        d->Nother < 0 || d->Nstc < 0 || d->Ndyn < 0)
        return MPCB_E_DIMS;
    prob_t P;
    P.d = d;
    P.rb = rb;
    P.p = p;
    P.L = make_layout(d);
    const layout_t* L = &P.L;
    const int N = d->N, n1 = 2 * N, n2 = n2_of(d);
    const double ts = rb->ts;
    const double* q = p + L->q;
    const double qN = q[5], qthN = q[6], accp = q[8], waccp = q[9];

    double st[MPCB_MAX_N + 1][3];
    st[0][0] = p[L->s_0];
    st[0][1] = p[L->s_0 + 1];
    st[0][2] = p[L->s_0 + 2];
    double cost = 0.0;
    double F2stack[256];
    double* F2 = n2 <= 256 ? F2stack : (double*)malloc((size_t)n2 * sizeof(double));
    memset(F2, 0, (size_t)n2 * sizeof(double));
    double F1[2 * MPCB_MAX_N];

    for (int k = 0; k < N; ++k) {
        rk4_step(st[k], u[2 * k], u[2 * k + 1], ts, st[k + 1]);
        stage_terms(&P, k, st[k + 1][0], st[k + 1][1], u[2 * k], u[2 * k + 1], 0, &cost,
                    F2, 0, 0.0, 0, 0);
    }
    /* terminal (mpc_builder.py:148) */
    {
        double ex = st[N][0] - p[L->s_N], ey = st[N][1] - p[L->s_N + 1];
        double et = st[N][2] - p[L->s_N + 2];
        cost += qN * (ex * ex + ey * ey) + qthN * (et * et);
    }
    /* accelerations (mpc_builder.py:156-169) */
    {
        double sa = 0.0, sw = 0.0;
        for (int k = 0; k < N; ++k) {
            double vp = k ? u[2 * (k - 1)] : p[L->u_m1];
            double wp = k ? u[2 * (k - 1) + 1] : p[L->u_m1 + 1];
            F1[k] = (u[2 * k] - vp) / ts;
            F1[N + k] = (u[2 * k + 1] - wp) / ts;
            sa += F1[k] * F1[k];
            sw += F1[N + k] * F1[N + k];
        }
        cost += sa * accp;
        cost += sw * waccp;
    }
    /* augmented part */
    double resid[2 * MPCB_MAX_N];
    double dist2 = 0.0, f2sq = 0.0;
    double cdiv = fmax(c, 1.0);
    for (int i = 0; i < n1; ++i) {
        double lo = i < N ? rb->lin_acc_min : -rb->ang_acc_max;
        double hi = i < N ? rb->lin_acc_max : rb->ang_acc_max;
        double z = F1[i] + (y ? y[i] : 0.0) / cdiv;
        resid[i] = rect_resid(z, lo, hi);
        dist2 += resid[i] * resid[i];
    }
    for (int i = 0; i < n2; ++i) f2sq += F2[i] * F2[i];
    double psi = cost + c * dist2 / 2.0 + c * f2sq / 2.0;

    if (f_out) *f_out = cost;
    if (psi_out) *psi_out = psi;
    if (F1_out) memcpy(F1_out, F1, sizeof(double) * (size_t)n1);
    if (F2_out) memcpy(F2_out, F2, sizeof(double) * (size_t)n2);

    if (grad) {
        /* reverse sweep: lam = d psi / d s_{k+1} */
        double lam[3] = {0.0, 0.0, 0.0};
        lam[0] = 2.0 * qN * (st[N][0] - p[L->s_N]);
        lam[1] = 2.0 * qN * (st[N][1] - p[L->s_N + 1]);
        lam[2] = 2.0 * qthN * (st[N][2] - p[L->s_N + 2]);
        for (int k = N - 1; k >= 0; --k) {
            double g[2], gu[2];
            double v = u[2 * k], w = u[2 * k + 1];
            stage_terms(&P, k, st[k + 1][0], st[k + 1][1], v, w, 1, 0, 0, F2, c, g, gu);
            lam[0] += g[0];
            lam[1] += g[1];
            double dx_dv, dy_dv, dx_dw, dy_dw, dx_dth, dy_dth;
            rk4_partials(st[k][2], v, w, ts, &dx_dv, &dy_dv, &dx_dw, &dy_dw, &dx_dth,
                         &dy_dth);
            grad[2 * k] = gu[0] + lam[0] * dx_dv + lam[1] * dy_dv;
            grad[2 * k + 1] = gu[1] + lam[0] * dx_dw + lam[1] * dy_dw + lam[2] * ts;
            lam[2] += lam[0] * dx_dth + lam[1] * dy_dth;
        }
        /* acceleration cost + ALM distance term: both act on F1 (tridiagonal) */
        for (int i = 0; i < n1; ++i) {
            int k = i < N ? i : i - N;
            int comp = i < N ? 0 : 1;
            double wq = (i < N ? accp : waccp);
            double gF = 2.0 * wq * F1[i] + c * resid[i]; /* d psi / d F1_i */
            grad[2 * k + comp] += gF / ts;
            if (k > 0) grad[2 * (k - 1) + comp] -= gF / ts;
        }
    }
    if (F2 != F2stack) free(F2);
    return MPCB_OK;
}

/* =========================================================================
 *  Solver half  [UPSTREAM restatement — parity unpinned]
 * ========================================================================= */

typedef struct {
    int n, mem;
    int active, first_old;
    double gamma;                 /* H0 scaling */
    double s[MPCB_MAX_LBFGS + 1][2 * MPCB_MAX_N];
    double yv[MPCB_MAX_LBFGS + 1][2 * MPCB_MAX_N];
    double rho[MPCB_MAX_LBFGS + 1];
    double alpha[MPCB_MAX_LBFGS + 1];
    double old_state[2 * MPCB_MAX_N], old_g[2 * MPCB_MAX_N];
    int order[MPCB_MAX_LBFGS + 1]; /* logical slot -> physical row (rotate_right) */
    double sy_eps, cbfgs_eps, cbfgs_alpha;
} lbfgs_t;

static double dot(const double* a, const double* b, int n)
{
    double s = 0.0;
    for (int i = 0; i < n; ++i) s += a[i] * b[i];
    return s;
}
static double norm2(const double* a, int n) { return sqrt(dot(a, a, n)); }

/* lbfgs crate: Lbfgs::reset */
static void lbfgs_reset(lbfgs_t* l)
{
    l->active = 0;
    l->first_old = 1;
}
static void lbfgs_init(lbfgs_t* l, int n, const mpcb_solver_cfg* cfg)
{
    l->n = n;
    l->mem = cfg->lbfgs_mem;
    l->gamma = 1.0;
    l->sy_eps = cfg->sy_epsilon;
    l->cbfgs_eps = cfg->cbfgs_epsilon;
    l->cbfgs_alpha = cfg->cbfgs_alpha;
    for (int i = 0; i <= l->mem; ++i) l->order[i] = i;
    lbfgs_reset(l);
}
/* lbfgs crate: Lbfgs::update_hessian(g, s) with new_s_and_y_valid */
static void lbfgs_update(lbfgs_t* l, const double* g, const double* state)
{
    int n = l->n, mem = l->mem;
    if (l->first_old) {
        l->first_old = 0;
        memcpy(l->old_state, state, sizeof(double) * (size_t)n);
        memcpy(l->old_g, g, sizeof(double) * (size_t)n);
        return;
    }
    int last = l->order[mem];
    double* sn = l->s[last];
    double* yn = l->yv[last];
    for (int i = 0; i < n; ++i) {
        sn[i] = state[i] - l->old_state[i];
        yn[i] = g[i] - l->old_g[i];
    }
    double ys = dot(sn, yn, n);
    double ss = dot(sn, sn, n);
    int ok;
    if (ss <= 2.2250738585072014e-308 || (l->sy_eps > 0.0 && ys <= l->sy_eps)) {
        ok = 0;
    } else if (l->cbfgs_eps > 0.0 && l->cbfgs_alpha > 0.0) {
        double lhs = ys / ss;
        double rhs = l->cbfgs_eps * pow(norm2(g, n), l->cbfgs_alpha);
        ok = lhs > rhs && isfinite(lhs) && isfinite(rhs);
    } else {
        ok = 1;
    }
    if (!ok) return; /* rejection: old_state/old_g are NOT advanced */
    memcpy(l->old_state, state, sizeof(double) * (size_t)n);
    memcpy(l->old_g, g, sizeof(double) * (size_t)n);
    /* rotate_right(1): the temp row becomes slot 0 */
    for (int i = mem; i > 0; --i) l->order[i] = l->order[i - 1];
    l->order[0] = last;
    for (int i = mem; i > 0; --i) l->rho[i] = l->rho[i - 1];
    l->rho[0] = 1.0 / ys;
    l->gamma = ys / dot(yn, yn, n);
    l->active = l->active + 1 < mem ? l->active + 1 : mem;
}
/* lbfgs crate: Lbfgs::apply_hessian (two-loop recursion, in place) */
static void lbfgs_apply(lbfgs_t* l, double* q)
{
    int n = l->n;
    if (l->active == 0) return;
    for (int k = 0; k < l->active; ++k) {
        const double* sk = l->s[l->order[k]];
        const double* yk = l->yv[l->order[k]];
        double a = l->rho[k] * dot(sk, q, n);
        l->alpha[k] = a;
        for (int i = 0; i < n; ++i) q[i] += -a * yk[i];
    }
    for (int i = 0; i < n; ++i) q[i] *= l->gamma;
    for (int k = l->active - 1; k >= 0; --k) {
        const double* sk = l->s[l->order[k]];
        const double* yk = l->yv[l->order[k]];
        double beta = l->rho[k] * dot(yk, q, n);
        double cf = l->alpha[k] - beta;
        for (int i = 0; i < n; ++i) q[i] += cf * sk[i];
    }
}

typedef struct {
    const mpcb_dims* d;
    const mpcb_robot* rb;
    const mpcb_solver_cfg* cfg;
    const double* p;
    int n;
    /* xi = (c, y) */
    double c;
    double y[2 * MPCB_MAX_N];
    /* PANOC cache */
    lbfgs_t lb;
    double grad[2 * MPCB_MAX_N], grad_prev[2 * MPCB_MAX_N], u_half[2 * MPCB_MAX_N],
        gfpr[2 * MPCB_MAX_N], dir[2 * MPCB_MAX_N], gstep[2 * MPCB_MAX_N],
        u_plus[2 * MPCB_MAX_N];
    double gamma, sigma, L, cost, norm_gfpr, tau, lhs_ls, rhs_ls, akkt_tol;
    int iter;
    long n_cost, n_grad, n_small;
} panoc_t;

/* ---- sensitivity switches (tests/ and scripts/solver_sensitivity.py only) ----------------------
 * The PANOC/ALM half restates upstream code that is absent from /root/reference and cannot be
 * built here.  The three details of that restatement most likely to differ from the real
 * optimization_engine are selectable, so that their effect on the returned status / solution
 * can be MEASURED instead of guessed (DESIGN.md "Oracle and parity status"):
 *   akkt  0: as restated - the previous gradient is cached at the top of step(), so the AKKT
 *            residual reduces to |gamma_fpr| / gamma
 *         1: no AKKT test at all (exit on |gamma_fpr| < tolerance alone)
 *         2: the previous gradient is the gradient at the PREVIOUS iterate (cached before the
 *            line search overwrites it): residual |gamma_fpr/gamma + grad(u_k) - grad(u_{k-1})|
 *   ls    0: as restated - an exhausted line search keeps its last trial point
 *         1: an exhausted line search falls back to the plain forward-backward step (tau = 0:
 *            u <- u_half, cost and gradient re-evaluated there)
 *   last  0: as restated - NotConvergedIterations whenever the outer-iteration cap was reached
 *         1: a solve whose exit criterion is met AT the last outer iteration reports the inner
 *            status (Converged)
 * Not thread-local: set before a batch, read by every worker. */
static int g_var_akkt = 0, g_var_ls = 0, g_var_last = 0;
void mpco_set_variant(int akkt, int ls, int last) { g_var_akkt = akkt; g_var_ls = ls; g_var_last = last; }

static double psi_cost(panoc_t* S, const double* u)
{
    double psi;
    mpco_eval(S->d, S->rb, S->p, u, S->y, S->c, 0, &psi, 0, 0, 0);
    S->n_cost++;
    return psi;
}
static void psi_grad(panoc_t* S, const double* u, double* g)
{
    mpco_eval(S->d, S->rb, S->p, u, S->y, S->c, 0, 0, g, 0, 0);
    S->n_grad++;
}
static void project_U(const panoc_t* S, double* u)
{
    for (int k = 0; k < S->n / 2; ++k) {
        double v = u[2 * k], w = u[2 * k + 1];
        /* Rectangle::project: clamp */
        u[2 * k] = v < S->rb->lin_vel_min ? S->rb->lin_vel_min
                                          : (v > S->rb->lin_vel_max ? S->rb->lin_vel_max : v);
        u[2 * k + 1] = w < -S->rb->ang_vel_max
                           ? -S->rb->ang_vel_max
                           : (w > S->rb->ang_vel_max ? S->rb->ang_vel_max : w);
    }
}
static void gradient_step(panoc_t* S, const double* u)
{
    for (int i = 0; i < S->n; ++i) S->gstep[i] = u[i] - S->gamma * S->grad[i];
}
static void half_step(panoc_t* S)
{
    memcpy(S->u_half, S->gstep, sizeof(double) * (size_t)S->n);
    project_U(S, S->u_half);
}
static void compute_fpr(panoc_t* S, const double* u)
{
    for (int i = 0; i < S->n; ++i) S->gfpr[i] = u[i] - S->u_half[i];
    S->norm_gfpr = norm2(S->gfpr, S->n);
}

/* PANOCEngine::init (+ LipschitzEstimator, delta 1e-12, epsilon 1e-6) */
static void panoc_init(panoc_t* S, double* u)
{
    int n = S->n;
    lbfgs_reset(&S->lb);
    S->lhs_ls = S->rhs_ls = 0.0;
    S->tau = 1.0;
    S->L = 0.0;
    S->sigma = 0.0;
    S->cost = 0.0;
    S->iter = 0;
    S->gamma = 0.0;
    S->cost = psi_cost(S, u);
    /* estimate_local_lipschitz: gradient at u, perturb u += h (left perturbed,
     * as upstream does), gradient at u+h */
    double h[2 * MPCB_MAX_N], gh[2 * MPCB_MAX_N];
    psi_grad(S, u, S->grad);
    for (int i = 0; i < n; ++i) h[i] = (1e-6 * u[i] > 1e-12) ? 1e-6 * u[i] : 1e-12;
    double norm_h = norm2(h, n);
    for (int i = 0; i < n; ++i) u[i] += h[i];
    psi_grad(S, u, gh);
    for (int i = 0; i < n; ++i) gh[i] -= S->grad[i];
    S->L = norm2(gh, n) / norm_h;
    S->gamma = 0.95 / fmax(S->L, 1e-10);
    S->sigma = (1.0 - 0.95) / (4.0 * S->gamma);
    gradient_step(S, u);
    half_step(S);
}

static int exit_condition(const panoc_t* S)
{
    if (!(S->norm_gfpr < S->cfg->tolerance)) return 0;
    if (g_var_akkt == 1) return 1;
    /* AKKT residual |gfpr/gamma + df - df_prev| (PANOCCache::akkt_residual) */
    double r = 0.0;
    for (int i = 0; i < S->n; ++i) {
        double t = S->gfpr[i] / S->gamma + S->grad[i] - S->grad_prev[i];
        r += t * t;
    }
    r = sqrt(r);
    return r < S->akkt_tol;
}

static double lipschitz_rhs(const panoc_t* S)
{
    double ip = dot(S->grad, S->gfpr, S->n);
    return S->cost + 1e-6 * fabs(S->cost) - ip +
           (0.95 / (2.0 * S->gamma)) * (S->norm_gfpr * S->norm_gfpr);
}

/* PANOCEngine::update_lipschitz_constant */
static void update_lipschitz(panoc_t* S, const double* u)
{
    double cost_half = psi_cost(S, S->u_half);
    S->cost = psi_cost(S, u);
    int it = 0;
    while (cost_half > lipschitz_rhs(S) && it < 10 && S->L < 1e9) {
        lbfgs_reset(&S->lb);
        S->L *= 2.0;
        S->gamma /= 2.0;
        gradient_step(S, u);
        half_step(S);
        cost_half = psi_cost(S, S->u_half);
        compute_fpr(S, u);
        it++;
    }
    S->sigma = (1.0 - 0.95) / (4.0 * S->gamma);
}

static double sqdiff(const double* a, const double* b, int n)
{
    double s = 0.0;
    for (int i = 0; i < n; ++i) s += (a[i] - b[i]) * (a[i] - b[i]);
    return s;
}

/* PANOCEngine::step ; returns 1 to continue */
static int panoc_step(panoc_t* S, double* u)
{
    int n = S->n;
    if (S->iter >= 1 && g_var_akkt != 2) memcpy(S->grad_prev, S->grad, sizeof(double) * (size_t)n);
    compute_fpr(S, u);
    if (exit_condition(S)) return 0;
    if (S->norm_gfpr < S->cfg->tolerance) S->n_small++;   /* AKKT test failed: the solve goes on */
    update_lipschitz(S, u);
    /* lbfgs_direction */
    lbfgs_update(&S->lb, S->gfpr, u);
    if (S->iter > 0) {
        memcpy(S->dir, S->gfpr, sizeof(double) * (size_t)n);
        lbfgs_apply(&S->lb, S->dir);
    }
    if (g_var_akkt == 2) memcpy(S->grad_prev, S->grad, sizeof(double) * (size_t)n);   /* grad(u_k), before the update */
    if (S->iter == 0) {
        /* update_no_linesearch */
        memcpy(u, S->u_half, sizeof(double) * (size_t)n);
        S->cost = psi_cost(S, u);
        psi_grad(S, u, S->grad);
        gradient_step(S, u);
        half_step(S);
    } else {
        /* linesearch on the forward-backward envelope */
        double dist2 = sqdiff(S->gstep, S->u_half, n);
        double fbe = S->cost - 0.5 * S->gamma * dot(S->grad, S->grad, n) +
                     0.5 * dist2 / S->gamma;
        S->rhs_ls = fbe - S->sigma * (S->norm_gfpr * S->norm_gfpr);
        S->tau = 1.0;
        int ls = 0;
        for (;;) {
            double om = 1.0 - S->tau;
            for (int i = 0; i < n; ++i)
                S->u_plus[i] = u[i] - om * S->gfpr[i] - S->tau * S->dir[i];
            S->cost = psi_cost(S, S->u_plus);
            psi_grad(S, S->u_plus, S->grad);
            for (int i = 0; i < n; ++i) S->gstep[i] = S->u_plus[i] - S->gamma * S->grad[i];
            half_step(S);
            double d2 = sqdiff(S->gstep, S->u_half, n);
            S->lhs_ls = S->cost - 0.5 * S->gamma * dot(S->grad, S->grad, n) +
                        0.5 * d2 / S->gamma;
            if (!(S->lhs_ls > S->rhs_ls && ls < 10)) break;
            S->tau /= 2.0;
            ls++;
        }
        /* upstream: on exhaustion sets tau=0 and u<-u_half, then overwrites
         * u<-u_plus unconditionally: the last trial point is what is kept. */
        if (g_var_ls == 1 && S->lhs_ls > S->rhs_ls) {
            /* variant: fall back to the forward-backward step u <- u_half of the iterate the search
             * started from (tau = 0: u - gfpr), cost and gradient re-evaluated there */
            for (int i = 0; i < n; ++i) S->u_plus[i] = u[i] - S->gfpr[i];
            S->cost = psi_cost(S, S->u_plus);
            psi_grad(S, S->u_plus, S->grad);
            for (int i = 0; i < n; ++i) S->gstep[i] = S->u_plus[i] - S->gamma * S->grad[i];
            half_step(S);
        }
        memcpy(u, S->u_plus, sizeof(double) * (size_t)n);
    }
    S->iter++;
    return 1;
}

/* PANOCOptimizer::solve ; returns inner exit status, iterations in *iters */
static int panoc_solve(panoc_t* S, double* u, int* iters, int iters_left /* <= 0: no budget */)
{
    panoc_init(S, u);
    int num_iter = 0, cont_it = 1, cont_rt = 1;
    int flag = panoc_step(S, u);
    while (flag && cont_it && cont_rt) {
        num_iter++;
        cont_it = num_iter < S->cfg->max_inner;
        /* continue_runtime: the reference's wall clock (max_solver_time) restated as a budget on
         * the inner iterations of the whole solve (cfg->max_inner_total) */
        cont_rt = iters_left <= 0 || num_iter < iters_left;
        flag = panoc_step(S, u);
    }
    *iters = num_iter;
    for (int i = 0; i < S->n; ++i)
        if (!isfinite(u[i])) return MPCB_NOT_FINITE_COMPUTATION;
    int status = !cont_it ? MPCB_NOT_CONVERGED_ITERATIONS
                          : (!cont_rt ? MPCB_NOT_CONVERGED_OUT_OF_TIME : MPCB_CONVERGED);
    memcpy(u, S->u_half, sizeof(double) * (size_t)S->n);
    return status;
}

/*
 * AlmOptimizer::solve for one instance.  out_scalars[11] =
 * {cost f(u*), last inner |gamma fpr|, f1_infeas, f2_norm, penalty c,
 *  n_outer, n_inner, n_cost_evals, n_grad_evals, exit_status,
 *  inner iterations with |gamma fpr| < tolerance that failed the AKKT test}.
 */
int32_t mpco_solve(const mpcb_dims* d, const mpcb_robot* rb, const mpcb_solver_cfg* cfg,
                   const double* p, const double* u0, const double* y0, const double* c0,
                   double* u_out, double* y_out, double* out_scalars)
{
    if (!d || !rb || !cfg || !p || !u_out) return MPCB_E_NULL;
    if (d->N < 1 || d->N > MPCB_MAX_N || cfg->lbfgs_mem < 1 ||
        cfg->lbfgs_mem > MPCB_MAX_LBFGS || d->nedge < 1 || d->nedge > MPCB_MAX_EDGE)
        return MPCB_E_DIMS;
    const int N = d->N, n = 2 * N, n1 = 2 * N, n2 = n2_of(d);
    panoc_t* S = (panoc_t*)calloc(1, sizeof(panoc_t));
    S->d = d;
    S->rb = rb;
    S->cfg = cfg;
    S->p = p;
    S->n = n;
    lbfgs_init(&S->lb, n, cfg);
    double u[2 * MPCB_MAX_N];
    for (int i = 0; i < n; ++i) u[i] = u0 ? u0[i] : 0.0;
    S->c = c0 ? *c0 : cfg->initial_penalty;
    for (int i = 0; i < n1; ++i) S->y[i] = y0 ? y0[i] : 0.0;
    S->akkt_tol = cfg->initial_tolerance;

    double y_plus[2 * MPCB_MAX_N] = {0.0}, F1[2 * MPCB_MAX_N];
    double* F2 = (double*)calloc((size_t)n2, sizeof(double));
    double dy = 0.0, dy_plus = 0.0, f2n = 0.0, f2n_plus = 0.0, last_fpr = -1.0;
    int alm_iter = 0, n_outer = 0, inner_total = 0;
    int exit_status = MPCB_CONVERGED;
    int failed = 0, met_criterion = 0;

    const int budget = cfg->max_inner_total > 0 ? cfg->max_inner_total : 0;
    for (int outer = 1; outer <= cfg->max_outer; ++outer) {
        /* AlmOptimizer::solve: "no time left" before an outer iteration -> NotConvergedOutOfTime */
        if (budget > 0 && inner_total >= budget) {
            exit_status = MPCB_NOT_CONVERGED_OUT_OF_TIME;
            break;
        }
        n_outer++;
        /* project y on Y = [-1e12,1e12]^n1 */
        for (int i = 0; i < n1; ++i)
            S->y[i] = S->y[i] < -1e12 ? -1e12 : (S->y[i] > 1e12 ? 1e12 : S->y[i]);
        /* set_akkt_tolerance zeroes the cached previous gradient */
        memset(S->grad_prev, 0, sizeof(S->grad_prev));
        int iters = 0;
        int inner_status = panoc_solve(S, u, &iters, budget > 0 ? budget - inner_total : 0);
        if (inner_status == MPCB_NOT_FINITE_COMPUTATION) {
            exit_status = MPCB_NOT_FINITE_COMPUTATION;
            failed = 1;
            break;
        }
        last_fpr = S->norm_gfpr;
        inner_total += iters;
        /* update_lagrange_multipliers: y+ = y + c (F1(u) - Proj_C(F1(u) + y/c)) */
        mpco_eval(d, rb, p, u, 0, 0.0, 0, 0, 0, F1, F2);
        for (int i = 0; i < n1; ++i) {
            double lo = i < N ? rb->lin_acc_min : -rb->ang_acc_max;
            double hi = i < N ? rb->lin_acc_max : rb->ang_acc_max;
            double z = F1[i] + S->y[i] / S->c;
            double pz = z < lo ? lo : (z > hi ? hi : z);
            y_plus[i] = S->y[i] + S->c * (F1[i] - pz);
        }
        f2n_plus = norm2(F2, n2);
        {
            double s = 0.0;
            for (int i = 0; i < n1; ++i) s += (y_plus[i] - S->y[i]) * (y_plus[i] - S->y[i]);
            dy_plus = sqrt(s);
        }
        /* is_exit_criterion_satisfied */
        const double EPS = 2.220446049250313e-16;
        int c1 = alm_iter > 0 && dy_plus <= S->c * cfg->delta_tolerance + EPS;
        int c2 = f2n_plus <= cfg->delta_tolerance + EPS;
        int c3 = S->akkt_tol <= cfg->tolerance + EPS;
        if (c1 && c2 && c3) {
            exit_status = inner_status;
            met_criterion = 1;
            break;
        }
        /* is_penalty_stall_criterion */
        int stall = alm_iter == 0 ||
                    (dy_plus <= cfg->sufficient_decrease * dy + EPS &&
                     f2n_plus <= cfg->sufficient_decrease * f2n + EPS);
        if (!stall) S->c *= cfg->penalty_update;
        S->akkt_tol = fmax(S->akkt_tol * cfg->inner_tol_update, cfg->tolerance);
        /* final_cache_update */
        alm_iter++;
        dy = dy_plus;
        f2n = f2n_plus;
        memcpy(S->y, y_plus, sizeof(double) * (size_t)n1);
        if (outer == cfg->max_outer) exit_status = MPCB_NOT_CONVERGED_ITERATIONS;
    }
    if (!failed && n_outer == cfg->max_outer && !(g_var_last == 1 && met_criterion)) exit_status = MPCB_NOT_CONVERGED_ITERATIONS;

    double cost = 0.0;
    mpco_eval(d, rb, p, u, 0, 0.0, &cost, 0, 0, 0, 0);
    memcpy(u_out, u, sizeof(double) * (size_t)n);
    if (y_out) memcpy(y_out, y_plus, sizeof(double) * (size_t)n1);
    if (out_scalars) {
        out_scalars[0] = cost;
        out_scalars[1] = last_fpr;
        out_scalars[2] = dy_plus / S->c;
        out_scalars[3] = f2n_plus;
        out_scalars[4] = S->c;
        out_scalars[5] = (double)n_outer;
        out_scalars[6] = (double)inner_total;
        out_scalars[7] = (double)S->n_cost;
        out_scalars[8] = (double)S->n_grad;
        out_scalars[9] = (double)exit_status;
        out_scalars[10] = (double)S->n_small;
    }
    free(F2);
    free(S);
    return MPCB_OK;
}

/* Batch driver used as the host-core CPU baseline (OpenMP over instances). */
int32_t mpco_solve_batch(const mpcb_dims* d, const mpcb_robot* rb,
                         const mpcb_solver_cfg* cfg, int32_t n_p, int32_t starts,
                         const double* p, const double* u0, double* u_out,
                         double* out_scalars /* [B,11] */, int32_t threads)
{
    const int np = make_layout(d).np, n = 2 * d->N;
    const long B = (long)n_p * starts;
    int rc = 0;
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads > 0 ? threads : 1)
#endif
    for (long b = 0; b < B; ++b) {
        int r = mpco_solve(d, rb, cfg, p + (b / starts) * np, u0 ? u0 + b * n : 0, 0, 0,
                           u_out + b * n, 0, out_scalars ? out_scalars + b * 11 : 0);
        if (r) rc = r;
    }
    return rc;
}

/* ---- primitive terms exported for the reference's known-answer vectors
 *      (src/tests/test_mpc_builder.py:16-253) ------------------------------- */
double mpco_dist_to_lineseg(double px, double py, double s1x, double s1y, double s2x,
                            double s2y)
{
    return sqrt(seg_dist_sq(px, py, s1x, s1y, s2x, s2y, 0, 0));
}
double mpco_inside_ellipse(double x, double y, double cx, double cy, double rx, double ry,
                           double ang)
{
    return ellipse_ind(x, y, cx, cy, rx, ry, ang, 0, 0);
}
double mpco_inside_cvx_polygon(double x, double y, const double* b, const double* a0,
                               const double* a1, int nedge)
{
    return polygon_ind(x, y, b, a0, a1, nedge, 0, 0);
}
void mpco_unicycle_rk4(const double* s, double v, double w, double ts, double* out)
{
    rk4_step(s, v, w, ts, out);
}
