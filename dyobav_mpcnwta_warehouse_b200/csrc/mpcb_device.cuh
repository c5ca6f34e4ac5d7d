// mpcb_device.cuh — device-side building blocks of the batched NMPC solver
// (sm_100a).  One warp owns one MPC instance; lane l owns horizon steps
// k = l + 32*j (j < SPL): its two decision variables (v_k, w_k) live in
// registers, the rollout is a warp prefix scan, the adjoint a suffix scan, all
// n=2N vector algebra of PANOC / L-BFGS is per-lane FMAs plus shuffle
// reductions.  The scenario (one parameter row p, shared by all multi-start
// guesses of that scenario) is a structure-of-arrays block in shared memory.
//
// What is computed follows the reference's problem definition
// (mpc_builder.py:45-174, mpc_cost.py, mpc_helper.py, motion_model.py:141-163)
// and OpEn's PANOC/ALM (see oracle/mpc_oracle.c for the restatement this is
// checked against).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

namespace mpcb {

constexpr unsigned FULL = 0xffffffffu;
constexpr int EF = 9;  // ellipse fields per slot
// ellipse field order inside a staged block
enum { E_CX = 0, E_CY, E_CA, E_SA, E_I1I, E_I2I, E_I1R, E_I2R, E_WAL };
// header slots
enum { H_S0X = 0, H_S0Y, H_S0T, H_UM1V, H_UM1W, H_SNX, H_SNY, H_SNT, H_Q = 8, H_SIZE = 20 };

struct Lay {            // offsets (in doubles) inside one staged scenario block
    int N, Nother, Nstc, nedge, Ndyn;
    int o_hdr, o_rv, o_qstc, o_seg, o_c0, o_c, o_poly, o_e0, o_et;
    int total;          // doubles, multiple of 2 (16-byte granularity for TMA bulk copies)
    int np;             // length of a raw parameter row
    // raw p offsets (mpc_builder.py:47-60)
    int p_um1, p_s0, p_sN, p_q, p_rs, p_rv, p_c0, p_c, p_os, p_od, p_qstc, p_qdyn;
};

struct KParams {
    Lay L;
    double ts, k6, inv_ts, ds2, vmargin, smargin;
    double vmin, vmax, wmax, amin, amax, wamax;
    double tol, tol0, delta, beta, rho, theta, c_init, sy_eps, cb_eps, cb_alpha;
    int max_inner, max_outer, mem;
    int n_p, starts, B;
    int warps, nsc;      // warps per CTA, scenario blocks per CTA
    int lb_doubles;      // per-warp L-BFGS scratch (doubles)
};

// ---------------------------------------------------------------- warp helpers
__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) v += __shfl_xor_sync(FULL, v, m);
    return v;
}
__device__ __forceinline__ double scan_incl(double v, int lane)
{
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        double t = __shfl_up_sync(FULL, v, d);
        if (lane >= d) v += t;
    }
    return v;
}
__device__ __forceinline__ double rscan_incl(double v, int lane)
{
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        double t = __shfl_down_sync(FULL, v, d);
        if (lane + d < 32) v += t;
    }
    return v;
}

// forward prefix over the (row j, lane) order; returns exclusive and inclusive sums
template <int SPL>
__device__ __forceinline__ void prefix(const double (&a)[SPL], double (&excl)[SPL],
                                       double (&incl)[SPL], int lane)
{
    double carry = 0.0;
#pragma unroll
    for (int j = 0; j < SPL; ++j) {
        double s = scan_incl(a[j], lane);
        double e = __shfl_up_sync(FULL, s, 1);
        if (lane == 0) e = 0.0;
        excl[j] = carry + e;
        incl[j] = carry + s;
        if (SPL > 1) carry += __shfl_sync(FULL, s, 31);
    }
}
// suffix sums: incl[j] = sum over steps >= own, excl[j] = sum over steps > own
template <int SPL>
__device__ __forceinline__ void suffix(const double (&a)[SPL], double (&excl)[SPL],
                                       double (&incl)[SPL], int lane)
{
    double carry = 0.0;
#pragma unroll
    for (int j = SPL - 1; j >= 0; --j) {
        double s = rscan_incl(a[j], lane);
        double e = __shfl_down_sync(FULL, s, 1);
        if (lane == 31) e = 0.0;
        excl[j] = carry + e;
        incl[j] = carry + s;
        if (SPL > 1) carry += __shfl_sync(FULL, s, 0);
    }
}
template <int SPL>
__device__ __forceinline__ double dotw(const double (&a0)[SPL], const double (&a1)[SPL],
                                       const double (&b0)[SPL], const double (&b1)[SPL])
{
    double s = 0.0;
#pragma unroll
    for (int j = 0; j < SPL; ++j) s = fma(a0[j], b0[j], fma(a1[j], b1[j], s));
    return warp_sum(s);
}

// ------------------------------------------------------------ primitive terms
struct EllT {  // everything one ellipse slot contributes
    double cost, gx, gy;     // weighted soft cost  w*alpha*max(0,E_infl)^2 and its gradient
    double hr, hrx, hry;     // raw hinge max(0,E_raw) and its gradient
};

// mpc_helper.py:38-52 / mpc_cost.py:26-44.  `f` points at field 0 of the slot,
// consecutive fields are `stride` doubles apart.
__device__ __forceinline__ void ellipse_terms(const bool GRAD, const double* __restrict__ f, int stride,
                                              double x, double y, EllT& o)
{
    const double ex = x - f[E_CX * stride], ey = y - f[E_CY * stride];
    const double ca = f[E_CA * stride], sa = f[E_SA * stride];
    const double a = fma(ex, ca, ey * sa);
    const double b = fma(ex, sa, -(ey * ca));
    const double a2 = a * a, b2 = b * b;
    const double i1 = f[E_I1I * stride], i2 = f[E_I2I * stride];
    const double Ei = fma(-b2, i2, fma(-a2, i1, 1.0));
    o.cost = 0.0; o.gx = 0.0; o.gy = 0.0; o.hr = 0.0; o.hrx = 0.0; o.hry = 0.0;
    if (Ei > 0.0) {   // inside the inflated ellipse (the raw one is contained in it)
        const double wal = f[E_WAL * stride];
        o.cost = wal * (Ei * Ei);
        if (GRAD) {
            const double ta = 2.0 * a * i1, tb = 2.0 * b * i2;
            const double dEx = -fma(ta, ca, tb * sa);
            const double dEy = -fma(ta, sa, -(tb * ca));
            const double m = 2.0 * wal * Ei;
            o.gx = m * dEx;
            o.gy = m * dEy;
        }
        const double r1 = f[E_I1R * stride], r2 = f[E_I2R * stride];
        const double Er = fma(-b2, r2, fma(-a2, r1, 1.0));
        if (Er > 0.0) {
            o.hr = Er;
            if (GRAD) {
                const double ta = 2.0 * a * r1, tb = 2.0 * b * r2;
                o.hrx = -fma(ta, ca, tb * sa);
                o.hry = -fma(ta, sa, -(tb * ca));
            }
        }
    }
}

// mpc_helper.py:54-75: I = prod_e max(0, b_e - a0_e x - a1_e y); rows {b,-a0,-a1}
__device__ __forceinline__ double polygon_ind(const bool GRAD, const double* __restrict__ e, int nedge,
                                              double x, double y, double& dIx, double& dIy)
{
    double I = 1.0;
    for (int j = 0; j < nedge; ++j) {
        const double r = fma(e[3 * j + 2], y, fma(e[3 * j + 1], x, e[3 * j]));
        I *= fmax(0.0, r);
    }
    dIx = 0.0; dIy = 0.0;
    if (GRAD && I > 0.0) {
        for (int j = 0; j < nedge; ++j) {
            double pr = 1.0;
            for (int m = 0; m < nedge; ++m)
                if (m != j) pr *= fma(e[3 * m + 2], y, fma(e[3 * m + 1], x, e[3 * m]));
            dIx = fma(pr, e[3 * j + 1], dIx);
            dIy = fma(pr, e[3 * j + 2], dIy);
        }
    }
    return I;
}

// ---------------------------------------------------------------- evaluation
template <int SPL>
struct EvalOut {
    double psi;          // augmented cost
    double f;            // plain cost (psi with c = 0)
    double f2sq;         // |F2|^2
    double gv[SPL], gw[SPL];
};

// Evaluate psi(u; c, y) (and its gradient) for the instance owned by this warp.
//   S: staged scenario block.  v/w: the point.  ya/yw: multipliers of the lane's
//   two F1 entries (acc_k, wacc_k).  F2out (nullable, global): per-obstacle F2.
template <int SPL>
__device__ __forceinline__ void eval_psi(const KParams& P, const double* __restrict__ S,
                                         const double (&v)[SPL], const double (&w)[SPL], double c,
                                         const double (&ya)[SPL], const double (&yw)[SPL],
                                         const bool GRAD, EvalOut<SPL>& out, int lane,
                                         double* F2out = nullptr)
{
    const Lay& L = P.L;
    const int N = L.N;
    const double* H = S + L.o_hdr;
    const double* q = H + H_Q;
    const double qvel = q[1], rv = q[3], rw = q[4], qN = q[5], qthN = q[6], qrpd = q[7];
    const double accp = q[8], waccp = q[9];

    bool act[SPL];
    int kk[SPL];
#pragma unroll
    for (int j = 0; j < SPL; ++j) {
        kk[j] = lane + 32 * j;
        act[j] = kk[j] < N;
    }

    // ---- rollout: theta by prefix sum, then RK4 increments, then positions
    double dth[SPL], th[SPL], thn[SPL];
#pragma unroll
    for (int j = 0; j < SPL; ++j) dth[j] = act[j] ? P.ts * w[j] : 0.0;
    prefix<SPL>(dth, th, thn, lane);
    double Cc[SPL], Ss[SPL], cb[SPL], sb[SPL], cc[SPL], sc[SPL], dx[SPL], dy[SPL];
#pragma unroll
    for (int j = 0; j < SPL; ++j) {
        const double t0 = H[H_S0T] + th[j];
        const double tb = t0 + 0.5 * dth[j];
        const double tc = H[H_S0T] + thn[j];
        double s0, c0;
        sincos(t0, &s0, &c0);
        sincos(tb, &sb[j], &cb[j]);
        sincos(tc, &sc[j], &cc[j]);
        Cc[j] = c0 + 4.0 * cb[j] + cc[j];
        Ss[j] = s0 + 4.0 * sb[j] + sc[j];
        const double kv = act[j] ? P.k6 * v[j] : 0.0;
        dx[j] = kv * Cc[j];
        dy[j] = kv * Ss[j];
    }
    double px[SPL], py[SPL], tmp[SPL];
    prefix<SPL>(dx, tmp, px, lane);
    prefix<SPL>(dy, tmp, py, lane);

    double cost = 0.0;
    double gx[SPL], gy[SPL], gvd[SPL], gwd[SPL];
    double Spoly[SPL], dSx[SPL], dSy[SPL];
    bool anyhinge = false;

#pragma unroll
    for (int j = 0; j < SPL; ++j) {
        const int k = act[j] ? kk[j] : N - 1;   // clamped index for loads
        const double x = H[H_S0X] + px[j], y = H[H_S0Y] + py[j];
        double cst = 0.0, ggx = 0.0, ggy = 0.0;

        // -- reference path: qrpd * min_{i>=k} dist^2(p, seg_i)   (mpc_cost.py:84-95)
        {
            const double* sg = S + L.o_seg;
            double best = INFINITY;
            int ib = k;
            for (int i = 0; i < N; ++i) {       // uniform loop: broadcast loads
                const double ex = x - sg[i], ey = y - sg[N + i];
                const double ddx = sg[2 * N + i], ddy = sg[3 * N + i];
                const double th_ = fma(ex, ddx, ey * ddy) * sg[4 * N + i];
                const double ts_ = fmin(fmax(th_, 0.0), 1.0);
                const double vx = fma(ts_, ddx, -ex), vy = fma(ts_, ddy, -ey);
                const double d2 = fma(vx, vx, vy * vy);
                if (i >= k && d2 < best) { best = d2; ib = i; }
            }
            cst = best * qrpd;
            if (GRAD) {
                const double ex = x - sg[ib], ey = y - sg[N + ib];
                const double ddx = sg[2 * N + ib], ddy = sg[3 * N + ib], inv = sg[4 * N + ib];
                const double th_ = fma(ex, ddx, ey * ddy) * inv;
                const double ts_ = fmin(fmax(th_, 0.0), 1.0);
                const double vx = fma(ts_, ddx, -ex), vy = fma(ts_, ddy, -ey);
                const double dt = (th_ > 0.0 && th_ < 1.0) ? 1.0 : ((th_ == 0.0 || th_ == 1.0) ? 0.5 : 0.0);
                const double vd = fma(vx, ddx, vy * ddy) * dt * inv;
                ggx = 2.0 * qrpd * fma(vd, ddx, -vx);
                ggy = 2.0 * qrpd * fma(vd, ddy, -vy);
            }
        }
        // -- speed reference + control effort (mpc_cost.py:46-53,78-79)
        {
            const double dv = v[j] - S[L.o_rv + k];
            cst = fma(qvel, dv * dv, cst);
            cst += rv * (v[j] * v[j]) + rw * (w[j] * w[j]);
            gvd[j] = 2.0 * qvel * dv + 2.0 * rv * v[j];
            gwd[j] = 2.0 * rw * w[j];
        }
        // -- fleet: linear hinge on squared distance (mpc_cost.py:65-76)
        {
            const double* c0x = S + L.o_c0;
            const double* c0y = c0x + L.Nother;
            double s1 = 0.0, s2 = 0.0;
            for (int r = 1; r < L.Nother; ++r) {     // robot 0 skipped (mpc_builder.py:86-87)
                const double ex = x - c0x[r], ey = y - c0y[r];
                const double h = P.ds2 - fma(ex, ex, ey * ey);
                if (h > 0.0) {
                    s1 += h;
                    if (GRAD) { ggx = fma(-2000.0, ex, ggx); ggy = fma(-2000.0, ey, ggy); }
                }
            }
            const double* cx_ = S + L.o_c;
            const double* cy_ = cx_ + L.Nother * N;
            for (int r = 0; r < L.Nother; ++r) {
                const double ex = x - cx_[r * N + k], ey = y - cy_[r * N + k];
                const double h = P.ds2 - fma(ex, ex, ey * ey);
                if (h > 0.0) {
                    s2 += h;
                    if (GRAD) { ggx = fma(-20.0, ex, ggx); ggy = fma(-20.0, ey, ggy); }
                }
            }
            cst = fma(1000.0, s1, cst);
            cst = fma(10.0, s2, cst);
        }
        // -- static polygons (mpc_builder.py:100-108)
        double sp = 0.0, spx = 0.0, spy = 0.0;
        {
            const double qs = S[L.o_qstc + k];
            const double* pe = S + L.o_poly;
            for (int i = 0; i < L.Nstc; ++i) {
                double dIx, dIy;
                const double I = polygon_ind(GRAD, pe + i * 3 * L.nedge, L.nedge, x, y, dIx, dIy);
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
        }
        // -- dynamic ellipses: t=0 slot (broadcast) and t=k+1 slot (mpc_builder.py:111-143)
        bool hinge = sp > 0.0;
        {
            const double* e0 = S + L.o_e0;
            const double* et = S + L.o_et + k;
            for (int i = 0; i < L.Ndyn; ++i) {
                EllT a, b;
                ellipse_terms(GRAD, e0 + i, L.Ndyn, x, y, a);
                ellipse_terms(GRAD, et + i * N, L.Ndyn * N, x, y, b);
                cst += a.cost + b.cost;
                if (GRAD) { ggx += a.gx + b.gx; ggy += a.gy + b.gy; }
                hinge |= (a.hr > 0.0) | (b.hr > 0.0);
            }
        }
        if (!act[j]) { cst = 0.0; ggx = 0.0; ggy = 0.0; sp = 0.0; spx = 0.0; spy = 0.0; hinge = false; gvd[j] = 0.0; gwd[j] = 0.0; }
        cost += cst;
        gx[j] = ggx; gy[j] = ggy;
        Spoly[j] = sp; dSx[j] = spx; dSy[j] = spy;
        anyhinge |= hinge;
    }

    // ---- terminal cost on the last state (mpc_builder.py:148)
    double gthN = 0.0;
    {
#pragma unroll
        for (int j = 0; j < SPL; ++j) {
            if (act[j] && kk[j] == N - 1) {
                const double ex = H[H_S0X] + px[j] - H[H_SNX], ey = H[H_S0Y] + py[j] - H[H_SNY];
                const double et = H[H_S0T] + thn[j] - H[H_SNT];
                cost += qN * fma(ex, ex, ey * ey) + qthN * (et * et);
                gx[j] = fma(2.0 * qN, ex, gx[j]);
                gy[j] = fma(2.0 * qN, ey, gy[j]);
                gthN = 2.0 * qthN * et;
            }
        }
    }

    // ---- penalty constraints F2 (mpc_builder.py:72,106,119,137; vector of Ndyn
    //      with the polygon hinge broadcast into every entry)
    double f2sq = 0.0;
    const int n2 = L.Ndyn > 0 ? L.Ndyn : 1;
    if (__any_sync(FULL, anyhinge) || F2out != nullptr) {
        double spl = 0.0;
#pragma unroll
        for (int j = 0; j < SPL; ++j) spl += Spoly[j];
        const double SP = warp_sum(spl);
        double sumF2 = 0.0;
        if (L.Ndyn == 0) {
            f2sq = SP * SP;
            sumF2 = SP;
            if (F2out && lane == 0) F2out[0] = SP;
        }
        for (int i = 0; i < L.Ndyn; ++i) {
            double hl = 0.0;
            EllT a[SPL], b[SPL];
#pragma unroll
            for (int j = 0; j < SPL; ++j) {
                const int k = act[j] ? kk[j] : N - 1;
                const double x = H[H_S0X] + px[j], y = H[H_S0Y] + py[j];
                ellipse_terms(GRAD, S + L.o_e0 + i, L.Ndyn, x, y, a[j]);
                ellipse_terms(GRAD, S + L.o_et + k + i * N, L.Ndyn * N, x, y, b[j]);
                if (!act[j]) { a[j].hr = 0.0; b[j].hr = 0.0; }
                hl += a[j].hr + b[j].hr;
            }
            double F2i = SP;
            if (__any_sync(FULL, hl > 0.0)) F2i += warp_sum(hl);
            if (F2out && lane == 0) F2out[i] = F2i;
            f2sq = fma(F2i, F2i, f2sq);
            sumF2 += F2i;
            if (GRAD && F2i > 0.0) {
                const double m = c * F2i;
#pragma unroll
                for (int j = 0; j < SPL; ++j) {
                    if (a[j].hr > 0.0) { gx[j] = fma(m, a[j].hrx, gx[j]); gy[j] = fma(m, a[j].hry, gy[j]); }
                    if (b[j].hr > 0.0) { gx[j] = fma(m, b[j].hrx, gx[j]); gy[j] = fma(m, b[j].hry, gy[j]); }
                }
            }
        }
        if (GRAD) {
            const double m = c * sumF2;
#pragma unroll
            for (int j = 0; j < SPL; ++j) {
                gx[j] = fma(m, dSx[j], gx[j]);
                gy[j] = fma(m, dSy[j], gy[j]);
            }
        }
    }
    (void)n2;

    // ---- accelerations: cost, ALM term on F1 (mpc_builder.py:156-169)
    double gFa[SPL], gFw[SPL];
    double dist2 = 0.0;
    {
        const double cdiv = fmax(c, 1.0);
        double vprev_carry = H[H_UM1V], wprev_carry = H[H_UM1W];
#pragma unroll
        for (int j = 0; j < SPL; ++j) {
            double vp = __shfl_up_sync(FULL, v[j], 1), wp = __shfl_up_sync(FULL, w[j], 1);
            if (lane == 0) { vp = vprev_carry; wp = wprev_carry; }
            if (SPL > 1) {
                vprev_carry = __shfl_sync(FULL, v[j], 31);
                wprev_carry = __shfl_sync(FULL, w[j], 31);
            }
            const double acc = (v[j] - vp) * P.inv_ts, wacc = (w[j] - wp) * P.inv_ts;
            const double za = acc + ya[j] / cdiv, zw = wacc + yw[j] / cdiv;
            const double ra = za > P.amax ? za - P.amax : (za < P.amin ? za - P.amin : 0.0);
            const double rwv = zw > P.wamax ? zw - P.wamax : (zw < -P.wamax ? zw + P.wamax : 0.0);
            if (act[j]) {
                cost = fma(accp, acc * acc, cost);
                cost = fma(waccp, wacc * wacc, cost);
                dist2 = fma(ra, ra, fma(rwv, rwv, dist2));
                gFa[j] = fma(2.0 * accp, acc, c * ra);
                gFw[j] = fma(2.0 * waccp, wacc, c * rwv);
            } else {
                gFa[j] = 0.0; gFw[j] = 0.0;
            }
        }
    }

    // ---- totals (butterfly: every lane ends with the same bits)
    const double f = warp_sum(cost);
    const double d2 = warp_sum(dist2);
    out.f = f;
    out.f2sq = f2sq;
    out.psi = f + 0.5 * c * d2 + 0.5 * c * f2sq;

    if (GRAD) {
        // adjoint: G = sum_{j>=k} g_j ; theta coupling via a second suffix sum
        double Gx[SPL], Gy[SPL], hh[SPL], Hex[SPL], Hin[SPL];
        const double gthN_all = warp_sum(gthN);   // only the lane owning step N-1 is non-zero
        suffix<SPL>(gx, tmp, Gx, lane);
        suffix<SPL>(gy, tmp, Gy, lane);
#pragma unroll
        for (int j = 0; j < SPL; ++j) hh[j] = fma(Gy[j], dx[j], -(Gx[j] * dy[j]));
        suffix<SPL>(hh, Hex, Hin, lane);
        double nFa_carry = 0.0, nFw_carry = 0.0;
#pragma unroll
        for (int j = SPL - 1; j >= 0; --j) {
            double na = __shfl_down_sync(FULL, gFa[j], 1), nw = __shfl_down_sync(FULL, gFw[j], 1);
            if (lane == 31) { na = nFa_carry; nw = nFw_carry; }
            if (SPL > 1) {
                nFa_carry = __shfl_sync(FULL, gFa[j], 0);
                nFw_carry = __shfl_sync(FULL, gFw[j], 0);
            }
            const double kv = P.k6 * v[j] * P.ts;
            const double dxw = -kv * fma(2.0, sb[j], sc[j]);
            const double dyw = kv * fma(2.0, cb[j], cc[j]);
            double g0 = gvd[j] + P.k6 * fma(Gx[j], Cc[j], Gy[j] * Ss[j]) + (gFa[j] - na) * P.inv_ts;
            double g1 = gwd[j] + fma(Gx[j], dxw, Gy[j] * dyw) + P.ts * (Hex[j] + gthN_all) +
                        (gFw[j] - nw) * P.inv_ts;
            out.gv[j] = act[j] ? g0 : 0.0;
            out.gw[j] = act[j] ? g1 : 0.0;
        }
    }
}

}  // namespace mpcb
