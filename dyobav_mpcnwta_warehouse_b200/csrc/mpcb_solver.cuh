// mpcb_solver.cuh — warp-resident PANOC + ALM/PM for one instance.
// Restates OpEn's optimization_engine (PANOCEngine::init/step,
// PANOCOptimizer::solve, AlmOptimizer::step/solve) and the lbfgs crate's
// update_hessian/apply_hessian; the reference configures them in
// mpc_builder.py:171-198.  oracle/mpc_oracle.c holds the CPU restatement this
// code is checked against, routine by routine.
#pragma once
#include "mpcb_device.cuh"
#include "../../include/mpcb.h"

namespace mpcb {

// L-BFGS memory of one warp, in shared memory: rows [M][2N] of s and y (lane k
// touches elements k and N+k of a row: conflict-free), rho[M], alpha[M].
template <int SPL>
struct Lbfgs {
    double* s;      // [M][2N]
    double* y;      // [M][2N]
    double* rho;    // [M]
    double* alpha;  // [M]
    int M, N, mem;
    int head;       // physical row of logical slot 0
    int active;
    bool first_old;
    double gamma;
    double os0[SPL], os1[SPL], og0[SPL], og1[SPL];   // old_state, old_g (registers)

    __device__ __forceinline__ void bind(double* base, int N_, int mem_)
    {
        N = N_; mem = mem_; M = mem_ + 1;
        s = base;
        y = s + M * 2 * N;
        rho = y + M * 2 * N;
        alpha = rho + M;
        head = 0; active = 0; first_old = true; gamma = 1.0;
    }
    __device__ __forceinline__ void reset() { active = 0; first_old = true; }
    __device__ __forceinline__ int phys(int logical) const
    {
        int r = head + logical;
        return r >= M ? r - M : r;
    }
};

template <int SPL>
struct Inst {   // per-lane slice of the solver state of one instance
    double u0[SPL], u1[SPL];       // iterate (v_k, w_k)
    double g0[SPL], g1[SPL];       // gradient_u
    // gradient_u_previous is not stored: at its only use (the AKKT residual at the top of a
    // step) it equals gradient_u from the second step on, and zero at the first
    double h0[SPL], h1[SPL];       // u_half_step
    double r0[SPL], r1[SPL];       // gamma_fpr
    double d0[SPL], d1[SPL];       // direction_lbfgs
    // gradient_step u - gamma*grad is recomputed where the line search needs it (same operands,
    // same bits) instead of being kept
    double ya[SPL], yw[SPL];       // Lagrange multipliers of (acc_k, wacc_k) / max(c, 1); y itself
                                   // lives in the warp's shared-memory scratch
    double gamma, sigma, Lc, cost, norm_r;
    int iter;
};

// Outer-loop (ALM) state of one instance: touched once per outer iteration, so it lives in the
// warp's shared-memory scratch for good.  Fields are read by every lane; every read-modify-write is
// done by lane 0 alone between two __syncwarp() (MPCB_CS_LANE0): under independent thread scheduling
// a lane that runs late (after a divergent walk, or lane 0 after waiting for the worker pool) must
// not read a value another lane has already incremented.
struct ColdState {
    double c;                      // penalty
    double akkt_tol;
    double dy, dy_plus, f2n, f2n_plus, last_fpr, fcost;
    int alm_iter, n_outer, inner_total, outer, inner, status, n_cost, n_grad;
    int failed, qscan, n_small, pad;
    unsigned long long t0, pad8;       // SPEC: %globaltimer when the instance was started.  sizeof == 128: keeps
                                       // the scratch a whole number of double2
};
// doubles of per-warp scratch after the L-BFGS rows: y, y+ (2N each), the parked solver vectors
// (7 x (v, w) per horizon step), the parked PANOC scalars and the cold state
__host__ __device__ constexpr int scratch_doubles(int N)
{
    return 4 * N + 7 * 2 * N + 16 + (int)(sizeof(ColdState) / 8);
}

#define MPCB_FORJ _Pragma("unroll") for (int j = 0; j < SPL; ++j)
#define MPCB_CS_LANE0(...) do { __syncwarp(); if (lane == 0) { __VA_ARGS__; } __syncwarp(); } while (0)

template <int SPL>
__device__ __forceinline__ double sumsq2(const double (&a)[SPL], const double (&b)[SPL])
{
    double s = 0.0;
    MPCB_FORJ s = fma(a[j], a[j], fma(b[j], b[j], s));
    return warp_sum(s);
}

template <int SPL>
__device__ __forceinline__ void project_U(const KParams& P, const double (&a0)[SPL],
                                          const double (&a1)[SPL], double (&o0)[SPL],
                                          double (&o1)[SPL])
{
    MPCB_FORJ {
        o0[j] = a0[j] < P.vmin ? P.vmin : (a0[j] > P.vmax ? P.vmax : a0[j]);
        o1[j] = a1[j] < -P.wmax ? -P.wmax : (a1[j] > P.wmax ? P.wmax : a1[j]);
    }
}

__device__ __forceinline__ unsigned long long globaltimer_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

template <int SPL>
__device__ __forceinline__ void grad_and_half_step(const KParams& P, Inst<SPL>& I,
                                                   const double (&p0)[SPL], const double (&p1)[SPL])
{
    double s0[SPL], s1[SPL];
    MPCB_FORJ {
        s0[j] = fma(-I.gamma, I.g0[j], p0[j]);
        s1[j] = fma(-I.gamma, I.g1[j], p1[j]);
    }
    project_U<SPL>(P, s0, s1, I.h0, I.h1);
}

// lbfgs crate: update_hessian(g = gamma_fpr, s = u)
template <int SPL>
__device__ __forceinline__ void lbfgs_update(const KParams& P, Lbfgs<SPL>& B, const Inst<SPL>& I,
                                             int lane, const bool (&act)[SPL])
{
    if (B.first_old) {
        B.first_old = false;
        MPCB_FORJ { B.os0[j] = I.u0[j]; B.os1[j] = I.u1[j]; B.og0[j] = I.r0[j]; B.og1[j] = I.r1[j]; }
        return;
    }
    double sn0[SPL], sn1[SPL], yn0[SPL], yn1[SPL];
    double ys = 0.0, ss = 0.0, yy = 0.0;
    MPCB_FORJ {
        sn0[j] = I.u0[j] - B.os0[j]; sn1[j] = I.u1[j] - B.os1[j];
        yn0[j] = I.r0[j] - B.og0[j]; yn1[j] = I.r1[j] - B.og1[j];
        ys = fma(sn0[j], yn0[j], fma(sn1[j], yn1[j], ys));
        ss = fma(sn0[j], sn0[j], fma(sn1[j], sn1[j], ss));
        yy = fma(yn0[j], yn0[j], fma(yn1[j], yn1[j], yy));
    }
    ys = warp_sum(ys); ss = warp_sum(ss); yy = warp_sum(yy);
    bool ok;
    if (ss <= 2.2250738585072014e-308 || (P.sy_eps > 0.0 && ys <= P.sy_eps)) {
        ok = false;
    } else if (P.cb_eps > 0.0 && P.cb_alpha > 0.0) {
        const double lhs = ddiv(ys, ss);
        const double rhs = P.cb_eps * (P.cb_alpha == 1.0 ? I.norm_r : pow(I.norm_r, P.cb_alpha));
        ok = lhs > rhs && isfinite(lhs) && isfinite(rhs);
    } else {
        ok = true;
    }
    if (!ok) return;
    MPCB_FORJ { B.os0[j] = I.u0[j]; B.os1[j] = I.u1[j]; B.og0[j] = I.r0[j]; B.og1[j] = I.r1[j]; }
    // rotate_right(1): the scratch row (logical M-1) becomes logical 0
    B.head = B.head == 0 ? B.M - 1 : B.head - 1;
    double* sr = B.s + B.head * 2 * B.N;
    double* yr = B.y + B.head * 2 * B.N;
    MPCB_FORJ {
        if (act[j]) {
            const int k = lane + 32 * j;
            sr[k] = sn0[j]; sr[B.N + k] = sn1[j];
            yr[k] = yn0[j]; yr[B.N + k] = yn1[j];
        }
    }
    if (lane == 0) B.rho[B.head] = ddiv(1.0, ys);
    B.gamma = ddiv(ys, yy);
    B.active = B.active + 1 < B.mem ? B.active + 1 : B.mem;
    __syncwarp();
}

// lbfgs crate: apply_hessian (two-loop recursion) on q = (d0, d1)
template <int SPL>
__device__ __forceinline__ void lbfgs_apply(Lbfgs<SPL>& B, double (&q0)[SPL], double (&q1)[SPL],
                                            int lane, const bool (&act)[SPL])
{
    if (B.active == 0) return;
#pragma unroll 1
    for (int k = 0; k < B.active; ++k) {
        const int row = B.phys(k);
        const double* sr = B.s + row * 2 * B.N;
        const double* yr = B.y + row * 2 * B.N;
        double sv0[SPL], sv1[SPL], yv0[SPL], yv1[SPL], part = 0.0;
        MPCB_FORJ {
            const int kk = act[j] ? lane + 32 * j : 0;
            sv0[j] = act[j] ? sr[kk] : 0.0; sv1[j] = act[j] ? sr[B.N + kk] : 0.0;
            yv0[j] = act[j] ? yr[kk] : 0.0; yv1[j] = act[j] ? yr[B.N + kk] : 0.0;
            part = fma(sv0[j], q0[j], fma(sv1[j], q1[j], part));
        }
        const double a = B.rho[row] * warp_sum(part);
        if (lane == 0) B.alpha[row] = a;
        MPCB_FORJ { q0[j] = fma(-a, yv0[j], q0[j]); q1[j] = fma(-a, yv1[j], q1[j]); }
    }
    __syncwarp();
    MPCB_FORJ { q0[j] *= B.gamma; q1[j] *= B.gamma; }
#pragma unroll 1
    for (int k = B.active - 1; k >= 0; --k) {
        const int row = B.phys(k);
        const double* sr = B.s + row * 2 * B.N;
        const double* yr = B.y + row * 2 * B.N;
        double sv0[SPL], sv1[SPL], part = 0.0;
        MPCB_FORJ {
            const int kk = act[j] ? lane + 32 * j : 0;
            sv0[j] = act[j] ? sr[kk] : 0.0; sv1[j] = act[j] ? sr[B.N + kk] : 0.0;
            const double y0 = act[j] ? yr[kk] : 0.0, y1 = act[j] ? yr[B.N + kk] : 0.0;
            part = fma(y0, q0[j], fma(y1, q1[j], part));
        }
        const double beta = B.rho[row] * warp_sum(part);
        const double cf = B.alpha[row] - beta;
        MPCB_FORJ { q0[j] = fma(cf, sv0[j], q0[j]); q1[j] = fma(cf, sv1[j], q1[j]); }
    }
}

// warp_sum through shared memory: every lane reads the 32 partials and adds them in the butterfly's own
// association ((i, i+16), then (i, i+8), ... - the tree every lane of the xor butterfly ends with), so the
// bits are the same; for a warp that is alone on its scheduler the 31 independent-ish adds are shorter than
// five dependent shuffle round trips.  (With 16 warps per SM the broadcast loads saturate the shared-memory
// crossbar: round 1 measured -43 % there; this is for the latency kernel's solving warp only.)
__device__ __forceinline__ double warp_sum_lds(double v, double* red, int lane)
{
    red[lane] = v;
    __syncwarp();
    const double2* r2 = reinterpret_cast<const double2*>(red);
    double a[32];
#pragma unroll
    for (int i = 0; i < 16; ++i) { const double2 t = r2[i]; a[2 * i] = t.x; a[2 * i + 1] = t.y; }
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = a[i] + a[i + 16];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = a[i] + a[i + 8];
#pragma unroll
    for (int i = 0; i < 4; ++i) a[i] = a[i] + a[i + 4];
    a[0] = a[0] + a[2]; a[1] = a[1] + a[3];
    __syncwarp();
    return a[0] + a[1];
}

// The same recursion with the rows of the next step loaded before the reduction of the current one (the
// latency kernel's lone solving warp has nothing else to cover the shared-memory latency with).  Same
// operations in the same order: same bits.
template <int SPL>
__device__ __forceinline__ void lbfgs_apply_prefetch(Lbfgs<SPL>& B, double (&q0)[SPL], double (&q1)[SPL],
                                                     int lane, const bool (&act)[SPL], double* red)
{
    if (B.active == 0) return;
    double sv0[SPL], sv1[SPL], yv0[SPL], yv1[SPL], rho;
    auto load = [&](int k) {
        const int row = B.phys(k);
        const double* sr = B.s + row * 2 * B.N;
        const double* yr = B.y + row * 2 * B.N;
        MPCB_FORJ {
            const int kk = act[j] ? lane + 32 * j : 0;
            sv0[j] = act[j] ? sr[kk] : 0.0; sv1[j] = act[j] ? sr[B.N + kk] : 0.0;
            yv0[j] = act[j] ? yr[kk] : 0.0; yv1[j] = act[j] ? yr[B.N + kk] : 0.0;
        }
        rho = B.rho[row];
    };
    load(0);
#pragma unroll 1
    for (int k = 0; k < B.active; ++k) {
        const int row = B.phys(k);
        double part = 0.0, c0[SPL], c1[SPL];
        MPCB_FORJ { part = fma(sv0[j], q0[j], fma(sv1[j], q1[j], part)); c0[j] = yv0[j]; c1[j] = yv1[j]; }
        const double rk = rho;
        if (k + 1 < B.active) load(k + 1);
        const double a = rk * warp_sum_lds(part, red, lane);
        if (lane == 0) B.alpha[row] = a;
        MPCB_FORJ { q0[j] = fma(-a, c0[j], q0[j]); q1[j] = fma(-a, c1[j], q1[j]); }
    }
    __syncwarp();
    MPCB_FORJ { q0[j] *= B.gamma; q1[j] *= B.gamma; }
    load(B.active - 1);
#pragma unroll 1
    for (int k = B.active - 1; k >= 0; --k) {
        const int row = B.phys(k);
        double part = 0.0, c0[SPL], c1[SPL];
        MPCB_FORJ { part = fma(yv0[j], q0[j], fma(yv1[j], q1[j], part)); c0[j] = sv0[j]; c1[j] = sv1[j]; }
        const double rk = rho, al = B.alpha[row];
        if (k > 0) load(k - 1);
        const double beta = rk * warp_sum_lds(part, red, lane);
        const double cf = al - beta;
        MPCB_FORJ { q0[j] = fma(cf, c0[j], q0[j]); q1[j] = fma(cf, c1[j], q1[j]); }
    }
}

template <int SPL>
__device__ __forceinline__ void compute_fpr(Inst<SPL>& I)
{
    MPCB_FORJ { I.r0[j] = I.u0[j] - I.h0[j]; I.r1[j] = I.u1[j] - I.h1[j]; }
    I.norm_r = dsqrt(sumsq2<SPL>(I.r0, I.r1));
}

struct SolveIO {
    const double* u0; const double* y0; const double* c0;
    double* u_out; double* cost; int32_t* exit_status; int32_t* n_outer; int32_t* n_inner;
    double* fpr; double* f1_infeas; double* f2_norm; double* penalty; double* y_out; int32_t* evals;
};

// ---------------------------------------------------------------- speculative line search
// LATENCY kernel for small batches (up to twelve instances per SM, two for N > 32; persistent CTAs, two
// resident per SM, one for N > 32): one CTA per instance at a time, warp 0 runs
// the solve as ever, SPEC_TRIALS helper warps evaluate the line-search trial points tau = 1, 1/2,
// 1/4, ... CONCURRENTLY instead of one after the other.  PANOC's line search accepts the first trial
// that passes the envelope test; on this problem it backtracks often (2.7 cost+gradient evaluations
// per iteration on average), and every evaluation is a dependent chain that one warp cannot overlap.
// Each trial is the same arithmetic whichever warp runs it, so results, statuses and the evaluation
// counters (only the trials up to the accepted one are counted) stay bit-identical to the one-warp
// kernel and to the laned oracle.
#ifndef MPCB_SPEC_TRIALS
#define MPCB_SPEC_TRIALS 4       // measured with the look-ahead warp (600 single solves; p50 / p95 ms): 3 trial
                                 // warps 25.2 / 66.9, 4: 23.9 / 61.8, 5: 25.1 / 64.9, 6: 25.0 / 61.6
#endif
// warp 0 solves, warps 1..TRIALS evaluate line-search trials, the last warp evaluates the look-ahead cost
constexpr int SPEC_THREADS = 32 * (2 + MPCB_SPEC_TRIALS);
constexpr int SPEC_TRIAL_THREADS = 32 * (1 + MPCB_SPEC_TRIALS);   // barriers 1, 2: solver + trial warps
constexpr int SPEC_CTAS_PER_SM = SPEC_THREADS <= 256 ? 2 : 1;     // what the register budget is cut for
// The named barriers of the latency kernel, out of line on purpose: the solving warp and the helper
// warps meet at ONE bar.sync instruction (the same address for every thread of the CTA), which is what
// compute-sanitizer's synccheck expects of the threads of a block.
#ifdef MPCB_SPEC_PROF
#define SPEC_T(i) do { const long long t_ = clock64(); spec_t[i] += t_ - spec_t0; spec_t0 = t_; } while (0)
#else
#define SPEC_T(i) do { } while (0)
#endif
// Barriers 1 / 2: a trial batch is published / evaluated (solver + trial warps); 3 / 4: a look-ahead
// request is published / served (solver + look-ahead warp).
template <int ID>
__device__ __noinline__ void spec_bar()
{
    asm volatile("bar.sync %0, %1;" ::"n"(ID), "n"(ID <= 2 ? SPEC_TRIAL_THREADS : 64) : "memory");
}
struct alignas(16) SpecShared {
    const double* S;
    double gamma, ceff;
    int cmd, ls0, par, la_cmd;     // par: which of the two result buffers the current batch fills
    int la_acc, la_par, pad0, pad1;   // look-ahead request: the half step of result (la_par, la_acc)
    double la_cost;                // ... and its answer
    double red[32];                // the solving warp's reduction scratch (warp_sum_lds)
    double la_ceff;                // the look-ahead warp's own copy of what is constant over an inner problem
    const double* la_S;            // (written while that warp is idle: the first line search of the problem)
    double pad2;
    double lhs[MPCB_SPEC_TRIALS], cost[MPCB_SPEC_TRIALS];
};
// doubles: header, request vectors [8][N] (u0,u1,r0,r1,d0,d1,ya,yw), results [2][TRIALS][6][N]
// (pt0,pt1,g0,g1,h0,h1; two buffers: the look-ahead warp reads the accepted half step of one batch while
// the next batch is being written)
__host__ __device__ constexpr int spec_doubles(int N)
{
    return (int)(sizeof(SpecShared) / 8) + 8 * N + 2 * MPCB_SPEC_TRIALS * 6 * N + 2 * N;   // + ya, yw of the look-ahead warp
}
static_assert(sizeof(SpecShared) % 16 == 0, "SpecShared keeps 16-byte granularity");

template <int SPL, int FIXED>
__device__ __forceinline__ void spec_helper(const KParams& P, SpecShared* SP, int h, int lane)
{
    const LayV<FIXED> LV{&P.L};
    const int N = LV.N();
    double* const req = reinterpret_cast<double*>(SP + 1);
    bool act[SPL];
    MPCB_FORJ act[j] = lane + 32 * j < N;
    for (;;) {
        spec_bar<1>();
        if (SP->cmd == 0) return;
        double* const res = req + 8 * N + (size_t)(SP->par * MPCB_SPEC_TRIALS + h) * 6 * N;
        const int lsj = SP->ls0 + h;
        if (lsj <= 10) {
            const double tau = ldexp(1.0, -lsj);
            const double gamma = SP->gamma;
            double u0[SPL], u1[SPL], pt0[SPL], pt1[SPL], ya[SPL], yw[SPL];
            MPCB_FORJ {
                const int k = act[j] ? lane + 32 * j : 0;
                u0[j] = act[j] ? req[k] : 0.0; u1[j] = act[j] ? req[N + k] : 0.0;
                const double r0 = act[j] ? req[2 * N + k] : 0.0, r1 = act[j] ? req[3 * N + k] : 0.0;
                const double d0 = act[j] ? req[4 * N + k] : 0.0, d1 = act[j] ? req[5 * N + k] : 0.0;
                ya[j] = act[j] ? req[6 * N + k] : 0.0; yw[j] = act[j] ? req[7 * N + k] : 0.0;
                pt0[j] = u0[j] - (1.0 - tau) * r0 - tau * d0;
                pt1[j] = u1[j] - (1.0 - tau) * r1 - tau * d1;
            }
            EvalOut<SPL> o;
            eval_psi<SPL, FIXED>(P, SP->S, pt0, pt1, SP->ceff, ya, yw, true, o, lane);
            __syncwarp();
            double h0[SPL], h1[SPL], s0[SPL], s1[SPL];
            MPCB_FORJ { s0[j] = fma(-gamma, o.gv[j], pt0[j]); s1[j] = fma(-gamma, o.gw[j], pt1[j]); }
            project_U<SPL>(P, s0, s1, h0, h1);
            double d2 = 0.0;
            MPCB_FORJ {
                const double e0 = fma(-gamma, o.gv[j], pt0[j]) - h0[j], e1 = fma(-gamma, o.gw[j], pt1[j]) - h1[j];
                d2 = fma(e0, e0, fma(e1, e1, d2));
            }
            d2 = warp_sum(d2);
            const double lhs = o.psi - 0.5 * gamma * sumsq2<SPL>(o.gv, o.gw) + div_nonneg(0.5 * d2, gamma);
            MPCB_FORJ {
                if (act[j]) {
                    const int k = lane + 32 * j;
                    res[k] = pt0[j]; res[N + k] = pt1[j]; res[2 * N + k] = o.gv[j]; res[3 * N + k] = o.gw[j];
                    res[4 * N + k] = h0[j]; res[5 * N + k] = h1[j];
                }
            }
            if (lane == 0) { SP->lhs[h] = lhs; SP->cost[h] = o.psi; }
        }
        __syncwarp();
        spec_bar<2>();
    }
}

// The look-ahead warp of the latency kernel: the next iteration opens with the cost at the accepted
// trial's half step (Lipschitz test); it is evaluated here while the solving warp updates its L-BFGS
// direction AND while the next trial batch runs - the test nearly always passes and nothing after it
// depends on that cost, so the solving warp looks at the answer one batch later (and, when the test
// fails after all, throws that batch away and takes the sequential path from the test on).
template <int SPL, int FIXED>
__device__ __forceinline__ void spec_lookahead(const KParams& P, SpecShared* SP, int lane)
{
    const LayV<FIXED> LV{&P.L};
    const int N = LV.N();
    const double* const req = reinterpret_cast<const double*>(SP + 1);
    bool act[SPL];
    MPCB_FORJ act[j] = lane + 32 * j < N;
    for (;;) {
        spec_bar<3>();
        if (SP->la_cmd == 0) return;
        const double* res = req + 8 * N + (size_t)(SP->la_par * MPCB_SPEC_TRIALS + SP->la_acc) * 6 * N;
        const double* lay = req + 8 * N + (size_t)2 * MPCB_SPEC_TRIALS * 6 * N;
        double h0[SPL], h1[SPL], ya[SPL], yw[SPL];
        MPCB_FORJ {
            const int k = act[j] ? lane + 32 * j : 0;
            h0[j] = act[j] ? res[4 * N + k] : 0.0; h1[j] = act[j] ? res[5 * N + k] : 0.0;
            ya[j] = act[j] ? lay[k] : 0.0; yw[j] = act[j] ? lay[N + k] : 0.0;
        }
        EvalOut<SPL> o;
        eval_psi<SPL, FIXED>(P, SP->la_S, h0, h1, SP->la_ceff, ya, yw, false, o, lane);
        __syncwarp();
        if (lane == 0) SP->la_cost = o.psi;
        __syncwarp();
        spec_bar<4>();
    }
}

// What the pending horizon evaluation is for.  The whole ALM/PANOC run is one
// loop around a single inlined eval_psi call site (small code footprint, no
// solver state in local memory): every handler ends by choosing the next point
// to evaluate.
enum EvalFor { ST_INIT, ST_INIT_LIP, ST_LIP_HALF, ST_LIP_U0, ST_LIP_LOOP, ST_NOLS, ST_LS, ST_ALM };

// AlmOptimizer::solve + PANOCOptimizer::solve + PANOCEngine::{init,step}.
//   MODE 0: solve the one instance b0 whose scenario block is S0, then return.
//   MODE 1: queue worker — pull instances from `counter` until the batch is exhausted.
//   TEAM: this warp is warp 0 of a team (one CTA per instance): the other warps evaluate the
//   ellipse cost terms of every horizon evaluation (mpcb_device.cuh "team mode").
//   SPEC: this warp is warp 0 of a latency CTA: helper warps evaluate the line-search trials concurrently.
template <int SPL, int MODE, int FIXED, bool TEAM = false, bool SPEC = false>
__device__ __forceinline__ void solve_worker(const KParams& P, const double* __restrict__ S0,
                                             const double* __restrict__ staged, int* __restrict__ counter,
                                             double* lb_mem, int b0, int lane, const SolveIO& io,
                                             TeamShared* T = nullptr, SpecShared* SP = nullptr)
{
    const LayV<FIXED> LV{&P.L};
    const int N = LV.N();
    bool act[SPL];
    MPCB_FORJ act[j] = lane + 32 * j < N;
    Inst<SPL> I;
    Lbfgs<SPL> B;
    B.bind(lb_mem, N, FIXED ? MPCB_FIX_MEM : P.mem);
    // y and y+ (2N doubles each) follow the L-BFGS rows; lane k only touches entries k and N+k
    double* const ysm = lb_mem + ((2 * B.M * 2 * N + 2 * B.M + 1) & ~1);
    double* const ypsm = ysm + 2 * N;
    // While the horizon evaluation runs (75 % of the time, ~100 live registers of its own) the
    // solver's vectors and most of its scalars are parked in shared memory: no spills inside the
    // evaluation's loops, and room for more resident warps per SM.
    double2* const vpark = reinterpret_cast<double2*>(ypsm + 2 * N);
    double* const spark = reinterpret_cast<double*>(vpark + 7 * N);
    ColdState* const CS = reinterpret_cast<ColdState*>(spark + 16);
#define MPCB_PARK_V(idx, a0, a1) MPCB_FORJ { if (act[j]) vpark[(idx) * N + lane + 32 * j] = make_double2(a0[j], a1[j]); }
#define MPCB_FILL_V(idx, a0, a1) MPCB_FORJ { if (act[j]) { const double2 t_ = vpark[(idx) * N + lane + 32 * j]; a0[j] = t_.x; a1[j] = t_.y; } }
    const double* __restrict__ S = S0;
    int b = b0;

    double pt0[SPL], pt1[SPL];          // the point of the pending evaluation
    double cost_half = 0.0, rhs_ls = 0.0, tau = 1.0, ceff = 0.0;
    int num_iter = 0, it_lip = 0, ls = 0;
    int st = ST_INIT;
    bool cont = true, flag = true, want_grad = true;
#ifdef MPCB_SPEC_PROF
    long long spec_t[8] = {0, 0, 0, 0, 0, 0, 0, 0}, spec_t0 = clock64();
#endif
    double spec_fbe = 0.0;      // SPEC: the envelope value of the accepted trial (= the next iteration's FBE
    bool spec_fbe_ok = false;   //       as long as gamma has not changed since)
    bool spec_la_in = false;    // SPEC: the look-ahead warp holds this inner problem's multipliers and penalty
    bool la_pending = false;    // SPEC: the cost at the half step is on its way (look-ahead warp)
    int spec_head0 = 0;         // SPEC: L-BFGS ring head and scaling before the update made ahead of the
    double spec_bg0 = 1.0;      //       Lipschitz test (restored if the test fails after all)
    int spec_par = 0;           // SPEC: result buffer of the current trial batch
    bool spec_dir = false;      // SPEC: the L-BFGS direction was computed ahead of the Lipschitz check
    const double EPS = 2.220446049250313e-16;
    EvalOut<SPL> o;
    if (lane == 0) CS->qscan = 0;
    __syncwarp();

L_fetch:
    if (MODE != 0) {
        // One queue per CTA: queue q owns scenarios q, q + nq, q + 2 nq, ... with all multi-start
        // guesses of a scenario adjacent, so the warps of a CTA work on two or three scenario
        // blocks at a time (L1 locality) instead of one each.  A warp whose own queue has run dry
        // steals single instances from the following queues.
        const int nq = gridDim.x;
        int sc = -1;
        for (int qs = CS->qscan; qs < nq; ++qs) {
            int q = (int)blockIdx.x + qs;
            if (q >= nq) q -= nq;
            const int nsc_q = q < P.n_p ? (P.n_p - q + nq - 1) / nq : 0;
            int t = 0;
            if (lane == 0) t = atomicAdd(counter + q, 1);
            t = __shfl_sync(FULL, t, 0);
            if (t < nsc_q * P.starts) {
                sc = q + (t / P.starts) * nq;
                if (P.order) sc = P.order[sc];
                b = sc * P.starts + t % P.starts;
                MPCB_CS_LANE0(CS->qscan = qs);
                break;
            }
        }
        if (sc < 0) {
            // launch profile: when this warp ran out of work (globaltimer, ns)
            if (lane == 0 && P.prof) {
                unsigned long long tns;
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tns));
                P.prof[MPCB_WS_PROF_CTAS + ((blockIdx.x * MPCB_WS_PROF_WARPS + (threadIdx.x >> 5)) & (MPCB_WS_PROF_CTAS * MPCB_WS_PROF_WARPS - 1))] = tns;
            }
            return;
        }
        S = staged + (size_t)sc * LV.total();
        // opaque from here on: under register pressure the compiler would otherwise re-derive the
        // pointer (an integer division) inside the evaluation instead of keeping it
        asm volatile("" : "+l"(S));
    }
    CS->n_cost = 0; CS->n_grad = 0; CS->n_small = 0;
    if constexpr (SPEC) { MPCB_CS_LANE0(CS->t0 = globaltimer_ns()); }
    MPCB_FORJ {
        const int k = lane + 32 * j;
        I.u0[j] = 0.0; I.u1[j] = 0.0; I.ya[j] = 0.0; I.yw[j] = 0.0;
        I.g0[j] = 0.0; I.g1[j] = 0.0;
        I.h0[j] = 0.0; I.h1[j] = 0.0; I.r0[j] = 0.0; I.r1[j] = 0.0;
        I.d0[j] = 0.0; I.d1[j] = 0.0;
        pt0[j] = 0.0; pt1[j] = 0.0;
        if (act[j]) {
            if (io.u0) {
                const double2 t = reinterpret_cast<const double2*>(io.u0 + (size_t)b * 2 * N)[k];
                I.u0[j] = t.x; I.u1[j] = t.y;
            }
            ysm[k] = io.y0 ? io.y0[(size_t)b * 2 * N + k] : 0.0;
            ysm[N + k] = io.y0 ? io.y0[(size_t)b * 2 * N + N + k] : 0.0;
            ypsm[k] = 0.0; ypsm[N + k] = 0.0;
        }
    }
    CS->dy = 0.0; CS->dy_plus = 0.0; CS->f2n = 0.0; CS->f2n_plus = 0.0; CS->last_fpr = -1.0; CS->fcost = 0.0;
    CS->alm_iter = 0; CS->n_outer = 0; CS->inner_total = 0; CS->outer = 1;
    CS->status = MPCB_CONVERGED; CS->failed = 0; CS->inner = MPCB_CONVERGED;
    CS->c = io.c0 ? io.c0[b] : P.c_init;
    CS->akkt_tol = P.tol0;
    I.gamma = 0.0; I.sigma = 0.0; I.Lc = 0.0; I.cost = 0.0; I.norm_r = 0.0; I.iter = 0;

L_outer_begin:   // ---- AlmOptimizer::step: project y on Y, then the inner problem
    __syncwarp();
    // AlmOptimizer::solve checks the remaining time before every outer iteration; here the clock
    // is the inner-iteration count (cfg->max_inner_total, 0 = off)
    if (P.budget > 0 && CS->inner_total >= P.budget) { CS->status = MPCB_NOT_CONVERGED_OUT_OF_TIME; goto L_finish; }
    if constexpr (SPEC) {
        // ... and, in the latency kernel, the reference's own rule: the wall clock (cfg->max_time_us)
        if (P.time_ns > 0) {
            const bool late = (long long)(__shfl_sync(FULL, globaltimer_ns(), 0) - CS->t0) >= P.time_ns;
            if (late) { CS->status = MPCB_NOT_CONVERGED_OUT_OF_TIME; goto L_finish; }
        }
    }
    MPCB_CS_LANE0(CS->n_outer = CS->n_outer + 1);
    MPCB_FORJ {
        const int k = lane + 32 * j;
        double ya_ = 0.0, yw_ = 0.0;
        if (act[j]) {
            ya_ = fmin(fmax(ysm[k], -1e12), 1e12);
            yw_ = fmin(fmax(ysm[N + k], -1e12), 1e12);
            ysm[k] = ya_; ysm[N + k] = yw_;
        }
        // psi uses y / max(c, 1): constant over the inner problem, divided once here
        I.ya[j] = ya_ / fmax(CS->c, 1.0);
        I.yw[j] = yw_ / fmax(CS->c, 1.0);
    }
    // PANOCEngine::init
    spec_fbe_ok = false;
    spec_la_in = false;
    B.reset();
    I.iter = 0; num_iter = 0; cont = true;
    MPCB_FORJ { pt0[j] = I.u0[j]; pt1[j] = I.u1[j]; }
    want_grad = true; ceff = CS->c; st = ST_INIT;

L_eval:
    // park
    MPCB_PARK_V(0, I.u0, I.u1); MPCB_PARK_V(1, I.g0, I.g1); MPCB_PARK_V(2, I.h0, I.h1);
    MPCB_PARK_V(3, I.r0, I.r1); MPCB_PARK_V(4, I.d0, I.d1);
    MPCB_PARK_V(5, B.os0, B.os1); MPCB_PARK_V(6, B.og0, B.og1);
    spark[0] = I.sigma; spark[1] = I.Lc; spark[2] = I.norm_r; spark[3] = cost_half;
    spark[4] = rhs_ls; spark[5] = tau; spark[6] = B.gamma;
    reinterpret_cast<int*>(spark + 8)[0] = num_iter;
    reinterpret_cast<int*>(spark + 8)[1] = it_lip;
    reinterpret_cast<int*>(spark + 8)[2] = ls;
    reinterpret_cast<int*>(spark + 8)[3] = B.head;
    reinterpret_cast<int*>(spark + 8)[4] = B.active;
    eval_psi<SPL, FIXED, TEAM>(P, S, pt0, pt1, ceff, I.ya, I.yw, want_grad, o, lane, nullptr, false, T);
    __syncwarp();   // reconverge after the divergent walks before the shared counters are touched
    // un-park (inactive lanes hold zeros in every vector)
    MPCB_FORJ {
        I.u0[j] = 0.0; I.u1[j] = 0.0; I.g0[j] = 0.0; I.g1[j] = 0.0;
        I.h0[j] = 0.0; I.h1[j] = 0.0; I.r0[j] = 0.0; I.r1[j] = 0.0; I.d0[j] = 0.0; I.d1[j] = 0.0;
        B.os0[j] = 0.0; B.os1[j] = 0.0; B.og0[j] = 0.0; B.og1[j] = 0.0;
    }
    MPCB_FILL_V(0, I.u0, I.u1); MPCB_FILL_V(1, I.g0, I.g1); MPCB_FILL_V(2, I.h0, I.h1);
    MPCB_FILL_V(3, I.r0, I.r1); MPCB_FILL_V(4, I.d0, I.d1);
    MPCB_FILL_V(5, B.os0, B.os1); MPCB_FILL_V(6, B.og0, B.og1);
    I.sigma = spark[0]; I.Lc = spark[1]; I.norm_r = spark[2]; cost_half = spark[3];
    rhs_ls = spark[4]; tau = spark[5]; B.gamma = spark[6];
    num_iter = reinterpret_cast<int*>(spark + 8)[0];
    it_lip = reinterpret_cast<int*>(spark + 8)[1];
    ls = reinterpret_cast<int*>(spark + 8)[2];
    B.head = reinterpret_cast<int*>(spark + 8)[3];
    B.active = reinterpret_cast<int*>(spark + 8)[4];
    MPCB_CS_LANE0(if (want_grad) CS->n_grad = CS->n_grad + 1; else CS->n_cost = CS->n_cost + 1);
    switch (st) {
        case ST_INIT: goto H_INIT;
        case ST_INIT_LIP: goto H_INIT_LIP;
        case ST_LIP_HALF: goto H_LIP_HALF;
        case ST_LIP_U0: goto H_LIP_U0;
        case ST_LIP_LOOP: goto H_LIP_LOOP;
        case ST_NOLS: goto H_NOLS;
        case ST_LS: goto H_LS;
        default: goto H_ALM;
    }

H_INIT: {   // cost and gradient at u; then perturb for the Lipschitz estimate
    I.cost = o.psi;
    double hs = 0.0;
    MPCB_FORJ {
        I.g0[j] = o.gv[j]; I.g1[j] = o.gw[j];
        // LipschitzEstimator: h = max(1e-6*u, 1e-12); u stays perturbed, as upstream
        const double h0 = act[j] ? ((1e-6 * I.u0[j] > 1e-12) ? 1e-6 * I.u0[j] : 1e-12) : 0.0;
        const double h1 = act[j] ? ((1e-6 * I.u1[j] > 1e-12) ? 1e-6 * I.u1[j] : 1e-12) : 0.0;
        hs = fma(h0, h0, fma(h1, h1, hs));
        I.u0[j] += h0; I.u1[j] += h1;
        pt0[j] = I.u0[j]; pt1[j] = I.u1[j];
    }
    I.norm_r = sqrt(warp_sum(hs));   // |h| parked here until H_INIT_LIP
    st = ST_INIT_LIP;
    goto L_eval;
}
H_INIT_LIP: {
    double t0[SPL], t1[SPL];
    MPCB_FORJ { t0[j] = o.gv[j] - I.g0[j]; t1[j] = o.gw[j] - I.g1[j]; }
    I.Lc = sqrt(sumsq2<SPL>(t0, t1)) / I.norm_r;
    I.gamma = 0.95 / fmax(I.Lc, 1e-10);
    I.sigma = (1.0 - 0.95) / (4.0 * I.gamma);
    grad_and_half_step<SPL>(P, I, I.u0, I.u1);
    goto L_step_begin;
}

L_step_begin:   // ---- PANOCEngine::step
    compute_fpr<SPL>(I);
    if (I.norm_r < P.tol) {
        double a = 0.0;
        MPCB_FORJ {
            // gradient_u_previous: a copy of gradient_u once a step has been taken, zero before
            const double t0 = div_maybe_zero(I.r0[j], I.gamma) + I.g0[j] - (I.iter >= 1 ? I.g0[j] : 0.0);
            const double t1 = div_maybe_zero(I.r1[j], I.gamma) + I.g1[j] - (I.iter >= 1 ? I.g1[j] : 0.0);
            a = fma(t0, t0, fma(t1, t1, a));
        }
        if (dsqrt(warp_sum(a)) < CS->akkt_tol) { flag = false; goto L_step_return; }
        // |gamma fpr| is below the tolerance but the AKKT residual |fpr| is not: the iteration goes on
        // (counted: on this workload most of the iterations of a non-converging solve are of this kind)
        MPCB_CS_LANE0(CS->n_small = CS->n_small + 1);
    }
    // update_lipschitz_constant: cost at the half step first
    if constexpr (SPEC) {
        if (la_pending) {
            // the look-ahead warp is evaluating it; the test nearly always passes and nothing after it
            // depends on that cost, so the L-BFGS update, the two-loop recursion and the next trial batch
            // go ahead now and the test is made when that batch comes back (H_LS part of this block)
            spec_head0 = B.head;
            spec_bg0 = B.gamma;
            SPEC_T(2);
            lbfgs_update<SPL>(P, B, I, lane, act);
            SPEC_T(6);
            MPCB_FORJ { I.d0[j] = I.r0[j]; I.d1[j] = I.r1[j]; }
            lbfgs_apply_prefetch<SPL>(B, I.d0, I.d1, lane, act, SP->red);
            SPEC_T(3);
            it_lip = 0;
            spec_dir = true;
            goto L_lip_check;
        }
    }
    MPCB_FORJ { pt0[j] = I.h0[j]; pt1[j] = I.h1[j]; }
    want_grad = false; st = ST_LIP_HALF;
    goto L_eval;

H_LIP_HALF:
    cost_half = o.psi;
    it_lip = 0;
    if (I.iter == 0) {   // cost at the (perturbed) start point; later iterations already hold it
        MPCB_FORJ { pt0[j] = I.u0[j]; pt1[j] = I.u1[j]; }
        st = ST_LIP_U0;
        goto L_eval;
    }
    goto L_lip_check;
H_LIP_U0:
    I.cost = o.psi;
    goto L_lip_check;
H_LIP_LOOP:
    cost_half = o.psi;
    compute_fpr<SPL>(I);
    ++it_lip;
    goto L_lip_check;

L_lip_check: {
    bool lip_fail = false;
    if (!(SPEC && spec_dir)) {   // (the latency kernel's look-ahead path has made this very test already: passed)
        const double ip = dotw<SPL>(I.g0, I.g1, I.r0, I.r1);
        const double rhs = I.cost + 1e-6 * fabs(I.cost) - ip + ddiv(0.95, 2.0 * I.gamma) * (I.norm_r * I.norm_r);
        lip_fail = cost_half > rhs && it_lip < 10 && I.Lc < 1e9;
    }
    if (lip_fail) {
        spec_fbe_ok = false;
        B.reset();
        I.Lc *= 2.0;
        I.gamma /= 2.0;
        grad_and_half_step<SPL>(P, I, I.u0, I.u1);
        MPCB_FORJ { pt0[j] = I.h0[j]; pt1[j] = I.h1[j]; }
        st = ST_LIP_LOOP;
        goto L_eval;
    }
    I.sigma = ddiv(1.0 - 0.95, 4.0 * I.gamma);
    // lbfgs_direction
    if (!(SPEC && spec_dir)) {
        lbfgs_update<SPL>(P, B, I, lane, act);
        if (I.iter > 0) {
            MPCB_FORJ { I.d0[j] = I.r0[j]; I.d1[j] = I.r1[j]; }
            lbfgs_apply<SPL>(B, I.d0, I.d1, lane, act);
        }
    }
    spec_dir = false;
    want_grad = true;
    if (I.iter == 0) {   // update_no_linesearch
        MPCB_FORJ { I.u0[j] = I.h0[j]; I.u1[j] = I.h1[j]; pt0[j] = I.u0[j]; pt1[j] = I.u1[j]; }
        st = ST_NOLS;
        goto L_eval;
    }
    // linesearch on the forward-backward envelope
    double fbe;
    if (SPEC && spec_fbe_ok) {
        // the helper that evaluated the accepted trial computed exactly this expression from exactly
        // these operands (its envelope test): taken over instead of two more reductions
        fbe = spec_fbe;
    } else {
        double dd = 0.0;
        MPCB_FORJ {   // gradient step of the current iterate: the operands it was last computed from
            const double e0 = fma(-I.gamma, I.g0[j], I.u0[j]) - I.h0[j], e1 = fma(-I.gamma, I.g1[j], I.u1[j]) - I.h1[j];
            dd = fma(e0, e0, fma(e1, e1, dd));
        }
        const double dist2 = warp_sum(dd);
        fbe = I.cost - 0.5 * I.gamma * sumsq2<SPL>(I.g0, I.g1) + div_nonneg(0.5 * dist2, I.gamma);
    }
    rhs_ls = fbe - I.sigma * (I.norm_r * I.norm_r);
    tau = 1.0;
    ls = 0;
    if constexpr (SPEC) {
        // hand the helper warps what a trial needs, then take the first trial (in the order
        // tau = 1, 1/2, ...) that passes, exactly as the sequential search would
        double* const req = reinterpret_cast<double*>(SP + 1);
        MPCB_FORJ {
            if (act[j]) {
                const int k = lane + 32 * j;
                req[k] = I.u0[j]; req[N + k] = I.u1[j]; req[2 * N + k] = I.r0[j]; req[3 * N + k] = I.r1[j];
                req[4 * N + k] = I.d0[j]; req[5 * N + k] = I.d1[j]; req[6 * N + k] = I.ya[j]; req[7 * N + k] = I.yw[j];
            }
        }
        if (lane == 0) { SP->S = S; SP->gamma = I.gamma; SP->ceff = CS->c; }
        if (!spec_la_in) {
            // first line search of an inner problem: the look-ahead warp is idle (no request pending)
            double* const lay = req + 8 * N + (size_t)2 * MPCB_SPEC_TRIALS * 6 * N;
            MPCB_FORJ {
                if (act[j]) { lay[lane + 32 * j] = I.ya[j]; lay[N + lane + 32 * j] = I.yw[j]; }
            }
            if (lane == 0) { SP->la_S = S; SP->la_ceff = CS->c; }
            spec_la_in = true;
        }
        for (;;) {
            spec_par ^= 1;
            if (lane == 0) { SP->ls0 = ls; SP->par = spec_par; SP->cmd = 1; }
            __syncwarp();
            SPEC_T(0);
            spec_bar<1>();
            spec_bar<2>();
            SPEC_T(1);
            if (la_pending) {
                // the Lipschitz test this step went past: cost at the half step from the look-ahead warp
                spec_bar<4>();
                SPEC_T(4);
                la_pending = false;
                cost_half = SP->la_cost;
                MPCB_CS_LANE0(CS->n_cost = CS->n_cost + 1);
                const double ip = dotw<SPL>(I.g0, I.g1, I.r0, I.r1);
                const double rhs = I.cost + 1e-6 * fabs(I.cost) - ip + ddiv(0.95, 2.0 * I.gamma) * (I.norm_r * I.norm_r);
                if (cost_half > rhs && I.Lc < 1e9) {
                    // it fails after all: the batch just evaluated is not part of the sequential run (not
                    // counted, not used); the L-BFGS buffer is reset by the failing branch, ring head and
                    // scaling are put back, and the step goes on from the test as the one-warp kernel would
                    B.head = spec_head0; B.gamma = spec_bg0;
#ifdef MPCB_SPEC_PROF
                    if (lane == 0 && P.prof) atomicAdd(P.prof + MPCB_WS_PROF_CTAS + 16008, 1ULL);   // how often this path runs
#endif
                    it_lip = 0;
                    want_grad = false;
                    spec_dir = false;
                    goto L_lip_check;
                }
            }
            int acc = -1;
#pragma unroll 1
            for (int t = 0; t < MPCB_SPEC_TRIALS && acc < 0 && ls + t <= 10; ++t) {
                MPCB_CS_LANE0(CS->n_grad = CS->n_grad + 1);
                if (!(SP->lhs[t] > rhs_ls && ls + t < 10)) acc = t;
            }
            if (acc >= 0) {
                const double* res = req + 8 * N + (size_t)(spec_par * MPCB_SPEC_TRIALS + acc) * 6 * N;
                I.cost = SP->cost[acc];
                MPCB_FORJ {
                    const int k = act[j] ? lane + 32 * j : 0;
                    I.u0[j] = act[j] ? res[k] : 0.0; I.u1[j] = act[j] ? res[N + k] : 0.0;
                    I.g0[j] = act[j] ? res[2 * N + k] : 0.0; I.g1[j] = act[j] ? res[3 * N + k] : 0.0;
                    I.h0[j] = act[j] ? res[4 * N + k] : 0.0; I.h1[j] = act[j] ? res[5 * N + k] : 0.0;
                }
                spec_fbe = SP->lhs[acc];
                spec_fbe_ok = true;
                // hand the accepted half step to the look-ahead warp
                __syncwarp();
                if (lane == 0) { SP->la_acc = acc; SP->la_par = spec_par; SP->la_cmd = 1; }
                __syncwarp();
                spec_bar<3>();
                la_pending = true;
                break;
            }
            ls += MPCB_SPEC_TRIALS;
        }
        I.iter++;
        flag = true;
        goto L_step_return;
    }
    MPCB_FORJ {
        pt0[j] = I.u0[j] - (1.0 - tau) * I.r0[j] - tau * I.d0[j];
        pt1[j] = I.u1[j] - (1.0 - tau) * I.r1[j] - tau * I.d1[j];
    }
    st = ST_LS;
    goto L_eval;
}
H_NOLS:
    I.cost = o.psi;
    MPCB_FORJ { I.g0[j] = o.gv[j]; I.g1[j] = o.gw[j]; }
    grad_and_half_step<SPL>(P, I, I.u0, I.u1);
    I.iter++;
    flag = true;
    goto L_step_return;
H_LS: {
    I.cost = o.psi;
    MPCB_FORJ { I.g0[j] = o.gv[j]; I.g1[j] = o.gw[j]; }
    grad_and_half_step<SPL>(P, I, pt0, pt1);
    double d2 = 0.0;
    MPCB_FORJ {
        const double e0 = fma(-I.gamma, I.g0[j], pt0[j]) - I.h0[j], e1 = fma(-I.gamma, I.g1[j], pt1[j]) - I.h1[j];
        d2 = fma(e0, e0, fma(e1, e1, d2));
    }
    d2 = warp_sum(d2);
    const double lhs = I.cost - 0.5 * I.gamma * sumsq2<SPL>(I.g0, I.g1) + div_nonneg(0.5 * d2, I.gamma);
    if (lhs > rhs_ls && ls < 10) {
        tau /= 2.0;
        ++ls;
        MPCB_FORJ {
            pt0[j] = I.u0[j] - (1.0 - tau) * I.r0[j] - tau * I.d0[j];
            pt1[j] = I.u1[j] - (1.0 - tau) * I.r1[j] - tau * I.d1[j];
        }
        goto L_eval;
    }
    // upstream keeps the last trial point even when the search is exhausted
    MPCB_FORJ { I.u0[j] = pt0[j]; I.u1[j] = pt1[j]; }
    I.iter++;
    flag = true;
    goto L_step_return;
}

L_step_return:   // PANOCOptimizer::solve: flag = step(); while (flag && cont) { ... step() }
    if (flag && cont) {
        ++num_iter;
        // continue_num_iters && continue_runtime (the runtime being the iteration budget)
        cont = num_iter < P.max_inner && (P.budget <= 0 || CS->inner_total + num_iter < P.budget);
        if constexpr (SPEC) {
            if (P.time_ns > 0) cont = cont && (long long)(__shfl_sync(FULL, globaltimer_ns(), 0) - CS->t0) < P.time_ns;
        }
        goto L_step_begin;
    }
    if constexpr (SPEC) {
        if (la_pending) { spec_bar<4>(); la_pending = false; }   // unused look-ahead
    }
    {
        bool fin = true;
        MPCB_FORJ fin = fin && isfinite(I.u0[j]) && isfinite(I.u1[j]);
        if (!__all_sync(FULL, fin)) { CS->status = MPCB_NOT_FINITE_COMPUTATION; CS->failed = 1; goto L_finish; }
    }
    MPCB_FORJ { I.u0[j] = I.h0[j]; I.u1[j] = I.h1[j]; }   // return u_bar (always feasible)
    __syncwarp();
    CS->inner = cont ? MPCB_CONVERGED
                     : (num_iter >= P.max_inner ? MPCB_NOT_CONVERGED_ITERATIONS : MPCB_NOT_CONVERGED_OUT_OF_TIME);
    CS->last_fpr = I.norm_r;
    MPCB_CS_LANE0(CS->inner_total = CS->inner_total + num_iter);
    // F1(u), F2(u), f(u) at the inner solution: one horizon evaluation with c = 0
    MPCB_FORJ { pt0[j] = I.u0[j]; pt1[j] = I.u1[j]; }
    want_grad = false; ceff = 0.0; st = ST_ALM;
    goto L_eval;

H_ALM: {
    __syncwarp();
    const double cpen = CS->c;
    const double fcost = o.f;
    const double f2n_plus = sqrt(o.f2sq);
    // update_lagrange_multipliers: y+ = y + c (F1 - Proj_C(F1 + y/c))
    double dsum = 0.0;
    double vc = S[LV.o_hdr() + H_UM1V], wc = S[LV.o_hdr() + H_UM1W];
    MPCB_FORJ {
        double vp = __shfl_up_sync(FULL, I.u0[j], 1), wp = __shfl_up_sync(FULL, I.u1[j], 1);
        if (lane == 0) { vp = vc; wp = wc; }
        if (SPL > 1) { vc = __shfl_sync(FULL, I.u0[j], 31); wc = __shfl_sync(FULL, I.u1[j], 31); }
        const int k = lane + 32 * j;
        const double ya_ = act[j] ? ysm[k] : 0.0, yw_ = act[j] ? ysm[N + k] : 0.0;
        const double acc = (I.u0[j] - vp) * P.inv_ts, wacc = (I.u1[j] - wp) * P.inv_ts;
        const double za = acc + ya_ / cpen, zw = wacc + yw_ / cpen;
        const double pa = fmin(fmax(za, P.amin), P.amax), pw = fmin(fmax(zw, -P.wamax), P.wamax);
        const double ypa = act[j] ? ya_ + cpen * (acc - pa) : 0.0;
        const double ypw = act[j] ? yw_ + cpen * (wacc - pw) : 0.0;
        if (act[j]) { ypsm[k] = ypa; ypsm[N + k] = ypw; }
        const double e0 = ypa - ya_, e1 = ypw - yw_;
        dsum = fma(e0, e0, fma(e1, e1, dsum));
    }
    const double dy_plus = sqrt(warp_sum(dsum));
    // every lane reads what it needs of the outer-loop state, then lane 0 alone updates it
    const int alm_iter = CS->alm_iter;
    const double akkt_tol = CS->akkt_tol;
    const double dy_old = CS->dy, f2n_old = CS->f2n;
    const int inner_st = CS->inner;
    // is_exit_criterion_satisfied
    const bool c1 = alm_iter > 0 && dy_plus <= cpen * P.delta + EPS;
    const bool c2 = f2n_plus <= P.delta + EPS;
    const bool c3 = akkt_tol <= P.tol + EPS;
    // is_penalty_stall_criterion
    const bool stall = alm_iter == 0 || (dy_plus <= P.theta * dy_old + EPS && f2n_plus <= P.theta * f2n_old + EPS);
    const bool exit_now = c1 && c2 && c3;
    MPCB_CS_LANE0(
        CS->fcost = fcost; CS->f2n_plus = f2n_plus; CS->dy_plus = dy_plus;
        if (exit_now) CS->status = inner_st;
        else {
            if (!stall) CS->c = cpen * P.rho;
            CS->akkt_tol = fmax(akkt_tol * P.beta, P.tol);
            CS->alm_iter = alm_iter + 1;
            CS->dy = dy_plus;
            CS->f2n = f2n_plus;
        });
    if (exit_now) goto L_finish;
    MPCB_FORJ {
        const int k = lane + 32 * j;
        if (act[j]) { ysm[k] = ypsm[k]; ysm[N + k] = ypsm[N + k]; }
    }
    {
        const bool more = CS->outer < P.max_outer;
        MPCB_CS_LANE0(if (more) CS->outer = CS->outer + 1);
        if (more) goto L_outer_begin;
    }
    goto L_finish;
}

L_finish:
    {
        const bool capped = !CS->failed && CS->n_outer == P.max_outer;
        MPCB_CS_LANE0(if (capped) CS->status = MPCB_NOT_CONVERGED_ITERATIONS);
    }
    MPCB_FORJ {
        const int k = lane + 32 * j;
        if (act[j]) {
            reinterpret_cast<double2*>(io.u_out + (size_t)b * 2 * N)[k] = make_double2(I.u0[j], I.u1[j]);
            if (io.y_out) {
                io.y_out[(size_t)b * 2 * N + k] = ypsm[k];
                io.y_out[(size_t)b * 2 * N + N + k] = ypsm[N + k];
            }
        }
    }
    if (lane == 0) {
        io.exit_status[b] = CS->status;
        if (io.cost) io.cost[b] = CS->failed ? __longlong_as_double(0x7ff8000000000000LL) : CS->fcost;
        if (io.n_outer) io.n_outer[b] = CS->n_outer;
        if (io.n_inner) io.n_inner[b] = CS->inner_total;
        if (io.fpr) io.fpr[b] = CS->last_fpr;
        if (io.f1_infeas) io.f1_infeas[b] = CS->dy_plus / CS->c;
        if (io.f2_norm) io.f2_norm[b] = CS->f2n_plus;
        if (io.penalty) io.penalty[b] = CS->c;
        if (io.evals) {
            io.evals[4 * b] = CS->n_cost; io.evals[4 * b + 1] = CS->n_grad;
            io.evals[4 * b + 2] = CS->n_small; io.evals[4 * b + 3] = 0;
        }
#ifdef MPCB_SPEC_PROF
        if (SPEC && P.prof) for (int i = 0; i < 8; ++i) P.prof[MPCB_WS_PROF_CTAS + 8 * (b & 127) + i] = (unsigned long long)spec_t[i];
#endif
    }
    __syncwarp();   // lane 0's output block is done before the next instance resets the counters
    if (MODE != 0) goto L_fetch;
}

}  // namespace mpcb
