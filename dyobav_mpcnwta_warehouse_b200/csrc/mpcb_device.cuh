// mpcb_device.cuh — device-side building blocks of the batched NMPC solver
// (sm_100a).  One warp owns one MPC instance; lane l owns horizon steps
// k = l + 32*j (j < SPL): its two decision variables (v_k, w_k) live in
// registers, the rollout is a warp prefix scan, the adjoint a suffix scan, all
// n=2N vector algebra of PANOC / L-BFGS is per-lane FMAs plus shuffle
// reductions.  The scenario (one parameter row p, shared by all multi-start
// guesses of that scenario) is a structure-of-arrays block (shared memory via
// TMA, or global memory through L1).
//
// What is computed follows the reference's problem definition
// (mpc_builder.py:45-174, mpc_cost.py, mpc_helper.py, motion_model.py:141-163)
// and OpEn's PANOC/ALM (see oracle/mpc_oracle.c for the restatement this is
// checked against).
//
// ARITHMETIC CONTRACT (mirrored operation-for-operation by the "laned" oracle,
// oracle/mpc_oracle_laned.c, which the GPU results must equal BIT FOR BIT):
//   * compiled with -fmad=false: only the fma() calls written here fuse;
//   * sums over the horizon are Kogge-Stone scans / xor-butterfly reductions in
//     the order written in scan_incl / rscan_incl / warp_sum;
//   * sin/cos come from sincos_cw below (Cody-Waite + fdlibm kernels), never
//     from the CUDA math library;
//   * obstacle terms that are provably zero (bounding-circle test around the
//     step's reference point) are skipped — adding an exact 0.0 does not change a
//     sum — and reference-path segments that provably cannot be the minimum are
//     not evaluated: with MPCB_CULL=0 every term is evaluated and the results
//     are bit-identical (tests/test_gpu_parity.py::test_culling_is_bit_exact).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

#ifndef MPCB_RED_ATTR
#define MPCB_RED_ATTR __noinline__
#endif
#ifndef MPCB_HELPER_ATTR
#define MPCB_HELPER_ATTR __forceinline__
#endif

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
    int o_mg;           // float margins start here (offset in doubles)
    // float sub-offsets (in floats) from o_mg.  Every table is [item][N]: entry (item, k) is a
    // conservative lower bound of the distance the robot must be from the anchor A_k = r_s[k]
    // (the reference point of step k) before that item can contribute at step k.
    int f_e0, f_et, f_poly, f_c0, f_c;
    int f_imin;         // per-item minimum over the steps: [Ndyn] ellipses (both slots), then
                        // [Nstc] polygons, [Nother] robots at t=0, [Nother] predicted robots
    int f_seg;          // [N][N]: (k, i) -> min over segments i' >= i of dist(A_k, segment i')
    int f_seg2;         // [N][N]: (k, i) -> min over segments i' >= i of (x - A_k).t_k, x on the segment,
                        // t_k the (slightly shortened) unit tangent of segment k: a directional bound
    int f_bx0;          // [Ndyn][4]: bounding box (xmin, xmax, ymin, ymax) of the inflated t = 0 ellipse
    int f_pbx;          // [Nstc][4]: bounding box (xmin, xmax, ymin, ymax) of a bounded polygon (outside it the
                        // polygon indicator is exactly zero); (-inf, inf, ...) for an unbounded one, empty for
                        // a zero-row (unused) slot
    int f_bx1;          // [Ndyn][NB][4]: bounding box of the inflated t = k+1 ellipses of the steps of block b
                        // (blocks of 8 steps, NB = ceil(N/8)); both 16-byte aligned, rounded outwards
    int total;          // doubles, multiple of 2 (16-byte granularity for TMA bulk copies)
    int np;             // length of a raw parameter row
    // raw p offsets (mpc_builder.py:47-60)
    int p_um1, p_s0, p_sN, p_q, p_rs, p_rv, p_c0, p_c, p_os, p_od, p_qstc, p_qdyn;
};

struct KParams {
    Lay L;
    double ts, k6, inv_ts, ds2, dsafe, vmargin, smargin;
    double vmin, vmax, wmax, amin, amax, wamax;
    double tol, tol0, delta, beta, rho, theta, c_init, sy_eps, cb_eps, cb_alpha;
    int max_inner, max_outer, mem;
    int n_p, starts, B;
    int warps, nsc;      // warps per CTA, scenario blocks per CTA (shared-memory variant)
    int lb_doubles;      // per-warp L-BFGS scratch (doubles)
    int cull;            // 1: skip provably-zero obstacle terms
    int budget;          // > 0: cap on the inner iterations of one solve (cfg->max_inner_total)
    long long time_ns;   // > 0: wall-clock cap on one solve (cfg->max_time_us; latency kernel only)
    unsigned long long* prof;   // launch profile in the workspace header (nullable): [CTAS] start, [CTAS][WARPS] finish
    const int* order;           // queue slot -> scenario, hardest first (nullable: identity)
    int team_G;                 // worker groups of the team kernels for this launch (0: one-warp kernels)
};
#include "../../include/mpcb.h"     // MPCB_WS_PROF_CTAS / MPCB_WS_PROF_WARPS

// One definition of the staged-block layout, usable at compile time (default dims are
// constant-folded into the kernel) and at run time (any dims).
__host__ __device__ constexpr Lay make_lay(int N, int Nother, int Nstc, int nedge, int Ndyn)
{
    Lay L{};
    L.N = N; L.Nother = Nother; L.Nstc = Nstc; L.nedge = nedge; L.Ndyn = Ndyn;
    int o = 0;
    L.o_hdr = o;  o += H_SIZE;
    L.o_rv = o;   o += N;
    L.o_qstc = o; o += N;
    L.o_seg = o;  o += 7 * N;   // s1x, s1y, dx, dy, 1/(|d|^2+1e-16), tangent x, y
    L.o_c0 = o;   o += 2 * Nother;
    L.o_c = o;    o += 2 * Nother * N;
    L.o_poly = o; o += 3 * nedge * Nstc;
    L.o_e0 = o;   o += EF * Ndyn;
    L.o_et = o;   o += EF * Ndyn * N;
    L.o_mg = o;
    int fo = 0;
    L.f_e0 = fo;   fo += Ndyn * N;
    L.f_et = fo;   fo += Ndyn * N;
    L.f_poly = fo; fo += Nstc * N;
    L.f_c0 = fo;   fo += Nother * N;
    L.f_c = fo;    fo += Nother * N;
    L.f_imin = fo; fo += Ndyn + Nstc + 2 * Nother;
    L.f_seg = fo;  fo += N * N;
    L.f_seg2 = fo; fo += N * N;
    while ((2 * L.o_mg + fo) & 3) ++fo;          // float4 loads
    L.f_bx0 = fo;  fo += 4 * Ndyn;
    L.f_bx1 = fo;  fo += 4 * Ndyn * ((N + 7) / 8);
    L.f_pbx = fo;  fo += 4 * Nstc;
    o += (fo + 1) / 2;
    L.total = (o + 1) & ~1;
    int q = 0;
    L.p_um1 = q; q += 2;
    L.p_s0 = q;  q += 3;
    L.p_sN = q;  q += 3;
    L.p_q = q;   q += 10;
    L.p_rs = q;  q += 3 * N;
    L.p_rv = q;  q += N;
    L.p_c0 = q;  q += 3 * Nother;
    L.p_c = q;   q += 3 * N * Nother;
    L.p_os = q;  q += 3 * nedge * Nstc;
    L.p_od = q;  q += 6 * (N + 1) * Ndyn;
    L.p_qstc = q; q += N;
    L.p_qdyn = q; q += N;
    L.np = q;
    return L;
}

// Dimension sets compiled into the solve kernel (FX > 0); FX = 0 is the run-time path.
//   1: config/mpc_default.yaml / mpc_fast.yaml (lines 21-31)      - BASELINE configs[0,1,3]
//   2: the same with 40 ellipses (2 pedestrians x 20 SWTA modes)  - BASELINE configs[2]
//   3: dense crowd, N = 40, 160 ellipses                          - BASELINE configs[4]
#define MPCB_FIX_MEM 10
__host__ __device__ constexpr Lay fixed_lay(int fx)
{
    return fx == 1 ? make_lay(20, 10, 10, 4, 15)
         : fx == 2 ? make_lay(20, 10, 10, 4, 40)
         : fx == 3 ? make_lay(40, 10, 10, 4, 160)
                   : make_lay(1, 0, 0, 1, 0);
}
constexpr int MPCB_NUM_FIXED = 3;

template <int FIXED>
struct LayV {
    const Lay* r;
#define MPCB_LAYF(name)                                                             \
    __device__ __forceinline__ int name() const                                     \
    {                                                                               \
        if constexpr (FIXED != 0) { constexpr int v = fixed_lay(FIXED).name; return v; } \
        else return r->name;                                                        \
    }
    MPCB_LAYF(N) MPCB_LAYF(Nother) MPCB_LAYF(Nstc) MPCB_LAYF(nedge) MPCB_LAYF(Ndyn)
    MPCB_LAYF(o_hdr) MPCB_LAYF(o_rv) MPCB_LAYF(o_qstc) MPCB_LAYF(o_seg) MPCB_LAYF(o_c0) MPCB_LAYF(o_c)
    MPCB_LAYF(o_poly) MPCB_LAYF(o_e0) MPCB_LAYF(o_et) MPCB_LAYF(o_mg)
    MPCB_LAYF(f_e0) MPCB_LAYF(f_et) MPCB_LAYF(f_poly) MPCB_LAYF(f_c0) MPCB_LAYF(f_c) MPCB_LAYF(f_imin) MPCB_LAYF(f_seg) MPCB_LAYF(f_seg2) MPCB_LAYF(f_bx0) MPCB_LAYF(f_bx1) MPCB_LAYF(f_pbx)
    MPCB_LAYF(total)
#undef MPCB_LAYF
};

// ---------------------------------------------------------------- team mode
// Dimension sets with many ellipses (Ndyn >= MPCB_TEAM_MIN_NDYN: the dense-crowd config, 160
// ellipses x 40 steps) are solved by CTAs of MPCB_TEAM_WARPS warps: MPCB_TEAM_SOLVERS "solver" warps,
// each running the solve of one instance exactly as the one-warp kernel does (rollout scans, adjoint
// scans, PANOC / L-BFGS algebra), share a pool of "worker" warps that evaluate everything of a
// horizon evaluation that is independent per step, in a flat (step, group) mapping: worker thread t
// owns step k = t % N of group g = t / N and handles the reference-path segments k+g, k+g+G, ...,
// the polygons and the ellipses i with i % G == g.  A solver publishes the positions of its pending
// evaluation and waits; the pool serves the solvers' requests one at a time, so the sequential part
// of one instance overlaps the parallel part of another.  Two or three instances per SM keep the
// part of their scenario blocks an evaluation touches resident in L1 (the one-warp kernel with 8
// instances per SM read it from DRAM every time: 22 stall cycles per issue on the long scoreboard);
// position-based culling (bounding boxes of the ellipses of a block of 8 steps against the box of
// the robot's positions in that block) replaces the anchor-based margins, which stop culling when
// the robot lags behind its reference.
// ARITHMETIC CONTRACT of team mode (mirrored by the laned oracle, G = team_groups(N, Ndyn)):
//   * the polygon and ellipse terms of step k (cost, position gradient, polygon hinge and its
//     gradient) are summed per group - polygons first, then ellipses, each in index order, from
//     +0.0 -, the G group sums are added together in group order (from +0.0), and that total is
//     added to the step's accumulators;
//   * the reference-path minimum is exact in any order (ties go to the lowest segment index);
//   * F2 keeps the one-warp order (F2_i = SP + butterfly over the lanes of the step hinges), and the
//     F2 share of the gradient of an ellipse with a raw hinge somewhere is
//     fx_k = fma(c F2_i, a.hrx + b.hrx, fx_k) for every step k (both slots added first).
#ifndef MPCB_TEAM_WARPS
#define MPCB_TEAM_WARPS 12
#endif
// which compiled-in dimension sets of the one-warp kernel use the position-based (block bounding
// box) culling of the ellipses instead of the anchor-based margins
#ifndef MPCB_BOX_CULL
#define MPCB_BOX_CULL(FIXED) ((FIXED) != 1)
#endif
// polygons: test the robot's position against the polygon's bounding box instead of the anchor margin
#ifndef MPCB_POLY_BOX
#define MPCB_POLY_BOX 1
#endif
#ifndef MPCB_TEAM_SOLVERS
#define MPCB_TEAM_SOLVERS 2
#endif
#ifndef MPCB_TEAM_MIN_NDYN
#define MPCB_TEAM_MIN_NDYN 64
#endif
#ifndef MPCB_TEAM_SLOTS
#define MPCB_TEAM_SLOTS 8        // hinge rows handed from the workers to a solver per evaluation; ellipses
                                 // with a hinge beyond that are redone by the solver warp (same values)
#endif
#ifndef MPCB_SPIN_SOLVER
#define MPCB_SPIN_SOLVER            // busy wait (a __nanosleep here costs far more than it saves)
#endif
#ifndef MPCB_SPIN_LEADER
#define MPCB_SPIN_LEADER
#endif
constexpr int TEAM_THREADS = 32 * MPCB_TEAM_WARPS;
constexpr int TEAM_NS = MPCB_TEAM_SOLVERS;
constexpr int TEAM_WORKERS = TEAM_THREADS - 32 * TEAM_NS;      // worker threads
#ifndef MPCB_TEAM_CTAS
#define MPCB_TEAM_CTAS (12 / MPCB_TEAM_WARPS)     // resident teams per SM the register budget is cut for
#endif
// number of worker groups (a power of two <= 32 with G*N <= 320: part of the arithmetic contract, so a
// constant and not the worker count of a particular build, which must be at least that); 0: one-warp kernel
#define MPCB_TEAM_MAP_THREADS 320
static_assert(TEAM_WORKERS >= MPCB_TEAM_MAP_THREADS, "the worker pool must cover the (step, group) mapping");
__host__ __device__ constexpr int team_groups(int N, int Ndyn, bool force = false)
{
    if (Ndyn < MPCB_TEAM_MIN_NDYN && !force) return 0;
    int g = 1;
    while (2 * g <= 32 && 2 * g * N <= MPCB_TEAM_MAP_THREADS) g *= 2;
    return g;
}
// Per-solver block: header, then X[N], Y[N] (request), SUM[6][N] (totals over the groups: cost, gx,
// gy, polygon hinge, its gradient), PATH[3][N] (path cost, gradient), HROW[SLOTS][N][3] (hinge, hrx,
// hry of the flagged ellipses) (results).
struct TeamShared {
    const double* S;       // scenario block of the instance being solved
    int grad;              // gradient wanted
    volatile int req;      // requests published by the solver ...
    volatile int done;     // ... and served by the pool
    volatile int exit_;    // the solver has no more instances
    int pad[6];
    unsigned hit[8];       // ellipses with a raw hinge at some step (set by the workers; Ndyn <= 256)
};
static_assert(sizeof(TeamShared) == 80, "TeamShared is 10 doubles");
struct TeamPtr { double *X, *Y, *SUM, *PATH, *HROW; };
__host__ __device__ constexpr int team_solver_doubles(int N)
{
    return (10 + 2 * N + 6 * N + 3 * N + 3 * MPCB_TEAM_SLOTS * N + 1) & ~1;   // even: 16-byte granularity
}
__device__ __forceinline__ TeamPtr team_ptrs(TeamShared* T, int N)
{
    TeamPtr q;
    q.X = reinterpret_cast<double*>(T) + 10;
    q.Y = q.X + N;
    q.SUM = q.Y + N;
    q.PATH = q.SUM + 6 * N;
    q.HROW = q.PATH + 3 * N;
    return q;
}
// Worker pool scratch: the solver being served, PART[6][G][N] (group sums), PB[G][N] / PI[G][N]
// (reference-path minimum and argmin per group).
struct TeamPool { volatile int cur; int pad[3]; };
__host__ __device__ constexpr int team_pool_doubles(int N, int G)
{
    return 2 + 6 * G * N + G * N + (G * N + 1) / 2;
}
// shared memory of a team CTA (doubles): per solver its warp scratch (lb_doubles) and its team
// block, then the pool
__host__ __device__ constexpr int team_smem_doubles(int N, int G, int lb_doubles, int solvers)
{
    return solvers * (lb_doubles + team_solver_doubles(N)) + team_pool_doubles(N, G);
}
__device__ __forceinline__ void bar_sync(int id, int nthreads)
{
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// ---------------------------------------------------------------- sin / cos
// Cody-Waite reduction by pi/2 (two constants, exact first product for
// |n| < 2^20) + the fdlibm __kernel_sin/__kernel_cos polynomials in Horner
// form.  |error| ~ 1 ulp for |x| < 1e5; identical code in the laned oracle.
#ifndef MPCB_SINCOS_ATTR
#define MPCB_SINCOS_ATTR __forceinline__
#endif
struct SinCos { double s, c; };
__host__ __device__ MPCB_SINCOS_ATTR SinCos sincos_cw_v(double x)
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
    const int q = static_cast<int>(fn) & 3;
    const double s1 = (q & 1) ? c : s;
    const double c1 = (q & 1) ? s : c;
    SinCos o;
    o.s = (q & 2) ? -s1 : s1;
    o.c = ((q + 1) & 2) ? -c1 : c1;
    return o;
}
__host__ __device__ __forceinline__ void sincos_cw(double x, double* sn, double* cs)
{
    const SinCos o = sincos_cw_v(x);
    *sn = o.s; *cs = o.c;
}

// ---------------------------------------------------------------- scalar helpers
// IEEE division / square root behind ONE out-of-line copy each: the solver's scalar bookkeeping
// divides in a dozen places, and every inlined a/b is ~25 instructions of Newton iteration plus
// a slow-path call.  The hot loop has to fit the SM's instruction cache (the kernel is bound by
// instruction supply, not by any pipe), so these are calls.  Same bits as the inline forms.
__device__ __noinline__ double ddiv(double a, double b) { return a / b; }
__device__ __noinline__ double dsqrt(double a) { return sqrt(a); }
// x / g for x >= 0: a zero numerator (no active bound: gradient step == half step) would send
// the inline division down its ~60-instruction slow path; 0 / g is +0 for any finite g > 0.
__device__ __forceinline__ double div_nonneg(double x, double g)
{
    if (x == 0.0 && g > 0.0 && g < INFINITY) return 0.0;
    return ddiv(x, g);
}

// x / g where x is often exactly zero (fixed-point residual of the steps past the horizon and of
// saturated inputs): zero lanes divide 1 by g instead and keep their (signed) zero, so the warp
// stays off the division's slow path; +-0 / g is +-0 for any finite g > 0.
__device__ __forceinline__ double div_maybe_zero(double x, double g)
{
    const bool z = x == 0.0 && g > 0.0 && g < INFINITY;
    const double q = ddiv(z ? 1.0 : x, g);
    return z ? x : q;
}

// max(0, r) and clamp to [0, 1] as plain selects (NaN -> 0, like fmax/fmin; a few instructions
// instead of the library forms' NaN/signed-zero handling).  Same code in the laned oracle.
__host__ __device__ __forceinline__ double pos_part(double r) { return r > 0.0 ? r : 0.0; }
__host__ __device__ __forceinline__ double clamp01(double t) { return t > 0.0 ? (t < 1.0 ? t : 1.0) : 0.0; }

// ---------------------------------------------------------------- warp helpers
__device__ MPCB_RED_ATTR double warp_sum(double v)
{
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) v += __shfl_xor_sync(FULL, v, m);
    return v;
}
// One Kogge-Stone stage on a double: the shuffle's own "source lane in range" predicate
// guards the add (2 SHFL + 1 predicated DADD per stage, no lane compare).
#define MPCB_SCAN_STAGE(DIR, CLAMP, V, D)                                                     \
    asm volatile("{\n\t"                                                                      \
                 ".reg .pred p;\n\t"                                                          \
                 ".reg .b32 lo, hi, tlo, thi;\n\t"                                            \
                 ".reg .f64 t;\n\t"                                                           \
                 "mov.b64 {lo, hi}, %0;\n\t"                                                  \
                 "shfl.sync." DIR ".b32 tlo|p, lo, %1, " CLAMP ", 0xffffffff;\n\t"             \
                 "shfl.sync." DIR ".b32 thi, hi, %1, " CLAMP ", 0xffffffff;\n\t"               \
                 "mov.b64 t, {tlo, thi};\n\t"                                                 \
                 "@p add.rn.f64 %0, %0, t;\n\t"                                               \
                 "}"                                                                          \
                 : "+d"(V)                                                                    \
                 : "r"(D))
__device__ MPCB_RED_ATTR double scan_incl(double v, int lane)
{
    (void)lane;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) MPCB_SCAN_STAGE("up", "0", v, d);
    return v;
}
__device__ MPCB_RED_ATTR double rscan_incl(double v, int lane)
{
    (void)lane;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) MPCB_SCAN_STAGE("down", "31", v, d);
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
template <int SPL>
__device__ __forceinline__ void prefix_incl(const double (&a)[SPL], double (&incl)[SPL], int lane)
{
    double carry = 0.0;
#pragma unroll
    for (int j = 0; j < SPL; ++j) {
        double s = scan_incl(a[j], lane);
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
__device__ __forceinline__ void suffix_incl(const double (&a)[SPL], double (&incl)[SPL], int lane)
{
    double carry = 0.0;
#pragma unroll
    for (int j = SPL - 1; j >= 0; --j) {
        double s = rscan_incl(a[j], lane);
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
__device__ MPCB_HELPER_ATTR void ellipse_terms(const bool GRAD, const double* __restrict__ f, int stride,
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
__device__ MPCB_HELPER_ATTR double polygon_ind(const bool GRAD, const double* __restrict__ e, int nedge,
                                              double x, double y, double& dIx, double& dIy)
{
    double I = 1.0;
#pragma unroll 1
    for (int j = 0; j < nedge; ++j) {
        const double r = fma(e[3 * j + 2], y, fma(e[3 * j + 1], x, e[3 * j]));
        I *= pos_part(r);
    }
    dIx = 0.0; dIy = 0.0;
    if (GRAD && I > 0.0) {
#pragma unroll 1
        for (int j = 0; j < nedge; ++j) {
            double pr = 1.0;
#pragma unroll 1
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
//   two F1 entries (acc_k, wacc_k) ALREADY DIVIDED by max(c, 1) — the quotient only changes
//   once per outer iteration, so the solver divides there and not in every evaluation.
//   F2out (nullable, global): per-obstacle F2.
#ifndef MPCB_EVAL_ATTR
#define MPCB_EVAL_ATTR __forceinline__
#endif
// -DMPCB_EVAL_PROF: clock64 segments of every evaluation, summed into the launch-profile area of the workspace
// (scripts/eval_prof.py): rollout | reference path | speed, control, fleet | polygons | ellipses + F2 |
// terminal, F2 gradient, accelerations, totals | adjoint; slot 7 counts evaluations
#ifdef MPCB_EVAL_PROF
#define EVAL_T(i) do { const long long t_ = clock64(); if (lane == 0 && P.prof) atomicAdd(P.prof + MPCB_WS_PROF_CTAS + 16000 + (i), (unsigned long long)(t_ - ev_t)); ev_t = clock64(); } while (0)
#else
#define EVAL_T(i) do { } while (0)
#endif
template <int SPL, int FIXED, bool TEAM = false>
__device__ MPCB_EVAL_ATTR void eval_psi(const KParams& P, const double* __restrict__ S,
                                         const double (&v)[SPL], const double (&w)[SPL], double c,
                                         const double (&ya)[SPL], const double (&yw)[SPL],
                                         const bool GRAD, EvalOut<SPL>& out, int lane,
                                         double* F2out = nullptr, const bool need_f = false,
                                         TeamShared* T = nullptr)
{
    const LayV<FIXED> L{&P.L};
    const int N = L.N();
    const double* H = S + L.o_hdr();
    const double* q = H + H_Q;
    const float* MG = reinterpret_cast<const float*>(S + L.o_mg());
    const bool CULL = P.cull != 0;

    bool act[SPL];
    int kk[SPL];
#pragma unroll
    for (int j = 0; j < SPL; ++j) {
        kk[j] = lane + 32 * j;
        act[j] = kk[j] < N;
    }

#ifdef MPCB_EVAL_PROF
    long long ev_t = clock64();
    if (lane == 0 && P.prof) atomicAdd(P.prof + MPCB_WS_PROF_CTAS + 16007, 1ULL);
#endif
    // ---- rollout: theta by prefix sum, then RK4 increments, then positions
    double dth[SPL], th[SPL], thn[SPL];
#pragma unroll
    for (int j = 0; j < SPL; ++j) dth[j] = act[j] ? P.ts * w[j] : 0.0;
    prefix<SPL>(dth, th, thn, lane);
    double c0s[SPL], s0s[SPL], cb[SPL], sb[SPL];
#pragma unroll
    for (int j = 0; j < SPL; ++j) {
        const double t0 = H[H_S0T] + th[j];
        const double tb = t0 + 0.5 * dth[j];
        sincos_cw(t0, &s0s[j], &c0s[j]);
        sincos_cw(tb, &sb[j], &cb[j]);
    }
    // cos/sin of the end-of-step heading = the next step's start heading: take
    // it from the next lane (the lane after the last step holds theta_N)
    double cc[SPL], sc[SPL], Cc[SPL], Ss[SPL], dx[SPL], dy[SPL];
#pragma unroll
    for (int j = 0; j < SPL; ++j) {
        double cn = __shfl_down_sync(FULL, c0s[j], 1), sn = __shfl_down_sync(FULL, s0s[j], 1);
        if (j + 1 < SPL) {
            const double c2 = __shfl_sync(FULL, c0s[j + 1 < SPL ? j + 1 : j], 0);
            const double s2 = __shfl_sync(FULL, s0s[j + 1 < SPL ? j + 1 : j], 0);
            if (lane == 31) { cn = c2; sn = s2; }
        } else if (N == 32 * SPL) {   // no lane after the last step: compute it
            if (lane == 31) sincos_cw(H[H_S0T] + thn[j], &sn, &cn);
        }
        cc[j] = cn; sc[j] = sn;
        Cc[j] = c0s[j] + 4.0 * cb[j] + cc[j];
        Ss[j] = s0s[j] + 4.0 * sb[j] + sc[j];
        const double kv = act[j] ? P.k6 * v[j] : 0.0;
        dx[j] = kv * Cc[j];
        dy[j] = kv * Ss[j];
    }
    double px[SPL], py[SPL];
    prefix_incl<SPL>(dx, px, lane);
    prefix_incl<SPL>(dy, py, lane);

    const double qvel = q[1], rv = q[3], rw = q[4], qrpd = q[7];
    double cost = 0.0;
    double gx[SPL], gy[SPL], gvd[SPL], gwd[SPL];
    double Spoly[SPL], dSx[SPL], dSy[SPL];
    double X[SPL], Y[SPL];
    float Df[SPL], Pf[SPL];
    double cstj[SPL];      // stage cost of the lane's step in row j
    const double* sg = S + L.o_seg();

    // positions and, per step, the distance to the step's anchor (its reference point),
    // inflated and rounded up: every culling test compares a precomputed lower bound with it
    float dmax = 0.f;
    {
        bool bad = false;
#pragma unroll
        for (int j = 0; j < SPL; ++j) {
            const int k = act[j] ? kk[j] : N - 1;
            X[j] = H[H_S0X] + px[j];
            Y[j] = H[H_S0Y] + py[j];
            const double ax = X[j] - sg[k], ay = Y[j] - sg[N + k];
            // upper bound of |p - A_k| in float: every step rounds up (the double square is inflated
            // past its own rounding first), so three instructions replace an IEEE double sqrt
            Df[j] = CULL ? __fsqrt_ru(__double2float_ru(fma(ax, ax, ay * ay) * (1.0 + 1e-9)))
                         : __int_as_float(0x7f800000);
            // how far the robot is AHEAD of the anchor along the path tangent, rounded up: a later
            // segment whose every point projects further ahead than that by more than the current
            // best distance cannot be the minimum (the lagging robot's walk stops after one segment)
            const double pr = fma(ax, sg[5 * N + k], ay * sg[6 * N + k]);
            Pf[j] = CULL ? __double2float_ru(fma(fabs(pr), 1e-9, pr) + 1e-9) : __int_as_float(0x7f800000);
            if (act[j]) {
                dmax = fmaxf(dmax, Df[j]);
                bad |= !(Df[j] < __int_as_float(0x7f800000));
            }
        }
#pragma unroll
        for (int m = 16; m > 0; m >>= 1) dmax = fmaxf(dmax, __shfl_xor_sync(FULL, dmax, m));
        if (__any_sync(FULL, bad)) dmax = __int_as_float(0x7f800000);   // non-finite state: cull nothing
    }
    if constexpr (TEAM) {
        // hand the positions to the worker pool and wait for its share of the evaluation
        const TeamPtr tp = team_ptrs(T, N);
#pragma unroll
        for (int j = 0; j < SPL; ++j)
            if (act[j]) { tp.X[kk[j]] = X[j]; tp.Y[kk[j]] = Y[j]; }
        if (lane < 8) T->hit[lane] = 0u;
        __syncwarp();
        // every lane waits in the same (warp-uniform) loop: a spin loop under `if (lane == 0)` leaves
        // the warp split into two convergence groups for the rest of the evaluation
        const int r = T->req + 1;
        __syncwarp();
        if (lane == 0) {
            T->S = S; T->grad = GRAD ? 1 : 0;
            __threadfence_block();
            T->req = r;
        }
        __syncwarp();
#ifdef MPCB_TEAM_BAR_WAIT
        // blocked on a named barrier the pool's first warp arrives at, instead of spinning on the flag: a
        // waiting solver warp then takes no issue slots from the workers of its scheduler
        bar_sync(8 + T->pad[0], 64);
#else
        while (T->done != r) { MPCB_SPIN_SOLVER; }
#endif
        __threadfence_block();
        __syncwarp();
    }
    const float* IM_E = MG + L.f_imin();
    const float* IM_P = IM_E + L.Ndyn();
    const float* IM_C0 = IM_P + L.Nstc();
    const float* IM_C = IM_C0 + L.Nother();

#pragma unroll
    for (int j = 0; j < SPL; ++j) {
        const int k = act[j] ? kk[j] : N - 1;   // clamped index for loads
        const double x = X[j], y = Y[j];
        const float D = Df[j];
        double cst = 0.0, ggx = 0.0, ggy = 0.0;
        EVAL_T(0);

        // -- reference path: qrpd * min_{i>=k} dist^2(p, seg_i)   (mpc_cost.py:84-95).
        //    Segments are visited from i = k; the walk stops once the precomputed bound
        //    says no later segment can beat the current minimum.
        if constexpr (TEAM) {       // the workers did it
            const TeamPtr tp = team_ptrs(T, N);
            cst = tp.PATH[k]; ggx = tp.PATH[N + k]; ggy = tp.PATH[2 * N + k];
        } else {
            const float* tm = MG + L.f_seg() + k * N;
            const float* tm2 = MG + L.f_seg2() + k * N;
            const float Pj = Pf[j];
            double best = INFINITY;
            int ib = k;
            bool alive = true;
#pragma unroll 1
            for (int i = k; __any_sync(FULL, alive); ++i) {
                if (alive) {
                    if (i >= N) { alive = false; }
                    else {
                        const float lb = fmaxf(tm[i] - D, tm2[i] - Pj);
                        if (lb > 0.f && (double)lb * (double)lb * (1.0 - 1e-6) > best) { alive = false; }
                        else {
                            const double ex = x - sg[i], ey = y - sg[N + i];
                            const double ddx = sg[2 * N + i], ddy = sg[3 * N + i];
                            const double th_ = fma(ex, ddx, ey * ddy) * sg[4 * N + i];
                            const double ts_ = clamp01(th_);
                            const double vx = fma(ts_, ddx, -ex), vy = fma(ts_, ddy, -ey);
                            const double d2 = fma(vx, vx, vy * vy);
                            if (d2 < best) { best = d2; ib = i; }
                        }
                    }
                }
            }
            cst = best * qrpd;
            if (GRAD) {
                const double ex = x - sg[ib], ey = y - sg[N + ib];
                const double ddx = sg[2 * N + ib], ddy = sg[3 * N + ib], inv = sg[4 * N + ib];
                const double th_ = fma(ex, ddx, ey * ddy) * inv;
                const double ts_ = clamp01(th_);
                const double vx = fma(ts_, ddx, -ex), vy = fma(ts_, ddy, -ey);
                const double dt = (th_ > 0.0 && th_ < 1.0) ? 1.0 : ((th_ == 0.0 || th_ == 1.0) ? 0.5 : 0.0);
                const double vd = fma(vx, ddx, vy * ddy) * dt * inv;
                ggx = 2.0 * qrpd * fma(vd, ddx, -vx);
                ggy = 2.0 * qrpd * fma(vd, ddy, -vy);
            }
        }
        EVAL_T(1);
        // -- speed reference + control effort (mpc_cost.py:46-53,78-79)
        {
            const double dv = v[j] - S[L.o_rv() + k];
            cst = fma(qvel, dv * dv, cst);
            cst += rv * (v[j] * v[j]) + rw * (w[j] * w[j]);
            gvd[j] = 2.0 * qvel * dv + 2.0 * rv * v[j];
            gwd[j] = 2.0 * rw * w[j];
        }
        // -- fleet: linear hinge on squared distance (mpc_cost.py:65-76).  A ballot over the
        //    per-item bounds yields the robots that can matter for ANY step of this evaluation;
        //    their bits are walked in index order, so sums keep the order of the full loop.
        {
            double s1 = 0.0, s2 = 0.0;
            const double* c0x = S + L.o_c0();
            const double* c0y = c0x + L.Nother();
            const float* m0 = MG + L.f_c0() + k;
            const double* cx_ = S + L.o_c();
            const double* cy_ = cx_ + L.Nother() * N;
            const float* m1 = MG + L.f_c() + k;
#pragma unroll 1
            for (int base = 0; base < L.Nother(); base += 32) {
                const int it = base + lane;
                unsigned ma = __ballot_sync(FULL, it >= 1 && it < L.Nother() && !(IM_C0[it < L.Nother() ? it : 0] > dmax));
                unsigned mb = __ballot_sync(FULL, it < L.Nother() && !(IM_C[it < L.Nother() ? it : 0] > dmax));
                while (ma) {                              // robot 0 skipped (mpc_builder.py:86-87)
                    const int r = base + __ffs(ma) - 1;
                    ma &= ma - 1;
                    if (m0[r * N] > D) continue;
                    const double ex = x - c0x[r], ey = y - c0y[r];
                    const double h = P.ds2 - fma(ex, ex, ey * ey);
                    if (h > 0.0) {
                        s1 += h;
                        if (GRAD) { ggx = fma(-2000.0, ex, ggx); ggy = fma(-2000.0, ey, ggy); }
                    }
                }
                while (mb) {
                    const int r = base + __ffs(mb) - 1;
                    mb &= mb - 1;
                    if (m1[r * N] > D) continue;
                    const double ex = x - cx_[r * N + k], ey = y - cy_[r * N + k];
                    const double h = P.ds2 - fma(ex, ex, ey * ey);
                    if (h > 0.0) {
                        s2 += h;
                        if (GRAD) { ggx = fma(-20.0, ex, ggx); ggy = fma(-20.0, ey, ggy); }
                    }
                }
            }
            cst = fma(1000.0, s1, cst);
            cst = fma(10.0, s2, cst);
        }
        EVAL_T(2);
        // -- static polygons (mpc_builder.py:100-108)
        double sp = 0.0, spx = 0.0, spy = 0.0;
        if constexpr (!TEAM) {
            const double qs = S[L.o_qstc() + k];
            const double* pe = S + L.o_poly();
            const float* mp = MG + L.f_poly() + k;
#pragma unroll 1
            for (int base = 0; base < L.Nstc(); base += 32) {
                const int it = base + lane;
                unsigned mk = __ballot_sync(FULL, it < L.Nstc() && !(IM_P[it < L.Nstc() ? it : 0] > dmax));
                while (mk) {
                    const int i = base + __ffs(mk) - 1;
                    mk &= mk - 1;
                    if (MPCB_POLY_BOX) {
                        // position test: outside the polygon's bounding box the indicator is exactly 0
                        const float4 qb = reinterpret_cast<const float4*>(MG + L.f_pbx())[i];
                        const float xf_lo = __double2float_rd(x), xf_hi = __double2float_ru(x);
                        const float yf_lo = __double2float_rd(y), yf_hi = __double2float_ru(y);
                        if (CULL && (xf_hi < qb.x || xf_lo > qb.y || yf_hi < qb.z || yf_lo > qb.w)) continue;
                    } else if (mp[i * N] > D) continue;
                    double dIx, dIy;
                    const double I = polygon_ind(GRAD, pe + i * 3 * L.nedge(), L.nedge(), x, y, dIx, dIy);
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
        }
        EVAL_T(3);
        // inactive lanes (steps beyond the horizon) carry no polygon hinge into F2
        cstj[j] = cst;
        gx[j] = ggx; gy[j] = ggy;
        Spoly[j] = act[j] ? sp : 0.0; dSx[j] = act[j] ? spx : 0.0; dSy[j] = act[j] ? spy : 0.0;
    }

    // ---- dynamic ellipses (mpc_builder.py:111-143) and the penalty constraints F2
    //      (mpc_builder.py:72,106,119,137: vector of Ndyn, the polygon hinge sum SP broadcast
    //      into every entry), in ONE walk over the candidate obstacles: every lane visits obstacle
    //      i at the same time, so F2_i = SP + sum over steps of the raw hinges is reduced right
    //      where the hinges are computed and the second evaluation pass of earlier versions is
    //      gone.  F2's gradient terms go to their own accumulator (fx, fy), added to the cost
    //      gradient once at the end, so the order of the cost-gradient sum does not depend on
    //      which obstacles are hit.
    if constexpr (TEAM) {
        // the workers' totals of the polygon and ellipse terms of every step
        const double* SUM = team_ptrs(T, N).SUM;
#pragma unroll
        for (int j = 0; j < SPL; ++j) {
            if (act[j]) {
                const double* q = SUM + kk[j];
                cstj[j] += q[0];
                Spoly[j] = q[3 * N];
                if (GRAD) { gx[j] += q[N]; gy[j] += q[2 * N]; dSx[j] = q[4 * N]; dSy[j] = q[5 * N]; }
            }
        }
    }
    double f2sq = 0.0, sumF2 = 0.0;
    double fx[SPL], fy[SPL];
    bool anyF2;     // whether any entry of F2 can be non-zero or is wanted
    {
        double spl = 0.0;
#pragma unroll
        for (int j = 0; j < SPL; ++j) { spl += Spoly[j]; fx[j] = 0.0; fy[j] = 0.0; }
        const bool anyp = __any_sync(FULL, spl > 0.0);
        const double SP = anyp ? warp_sum(spl) : 0.0;
        anyF2 = anyp || F2out != nullptr;
        int nxt = 0;     // F2 entries [0, nxt) are accounted for
        int nflag = 0;   // team mode: flagged ellipses met so far (= the row the workers filled)
        const double* e0 = S + L.o_e0();
        const double* etb = S + L.o_et();
        const float* me0b = MG + L.f_e0();
        const float* metb = MG + L.f_et();
        // Position-based culling (BOX): the bounding box of the robot's positions over a block of 8
        // steps against the bounding boxes of the ellipses of that block (staged by K3).  Unlike the
        // anchor-based margins it keeps culling when the robot lags behind its reference.  Exact in
        // the same sense: a skipped term is provably zero.
        constexpr bool BOX = !TEAM && MPCB_BOX_CULL(FIXED);
        float rlx[SPL], rhx[SPL], rly[SPL], rhy[SPL];
        if constexpr (BOX) {
            const float PINF = __int_as_float(0x7f800000);
            const bool nocull = !(dmax < PINF);          // culling off, or a non-finite state
#pragma unroll
            for (int j = 0; j < SPL; ++j) {
                rlx[j] = act[j] ? __double2float_rd(X[j]) : PINF; rhx[j] = act[j] ? __double2float_ru(X[j]) : -PINF;
                rly[j] = act[j] ? __double2float_rd(Y[j]) : PINF; rhy[j] = act[j] ? __double2float_ru(Y[j]) : -PINF;
#pragma unroll
                for (int m = 1; m < 8; m <<= 1) {
                    rlx[j] = fminf(rlx[j], __shfl_xor_sync(FULL, rlx[j], m)); rhx[j] = fmaxf(rhx[j], __shfl_xor_sync(FULL, rhx[j], m));
                    rly[j] = fminf(rly[j], __shfl_xor_sync(FULL, rly[j], m)); rhy[j] = fmaxf(rhy[j], __shfl_xor_sync(FULL, rhy[j], m));
                }
                if (nocull) { rlx[j] = -PINF; rhx[j] = PINF; rly[j] = -PINF; rhy[j] = PINF; }
            }
        }
#pragma unroll 1
        for (int base = 0; base < L.Ndyn(); base += 32) {
            const int it = base + lane;
            unsigned my0[SPL], my1[SPL];   // BOX: ellipses whose t = 0 / t = k+1 box meets the box of this lane's block
            unsigned mk;
            if constexpr (TEAM) {
                // team mode: only the ellipses a worker flagged (a raw hinge somewhere) are redone here
                mk = T->hit[base >> 5];
            } else if constexpr (BOX) {
                const bool valid = it < L.Ndyn();
                const int o = valid ? it : 0;
                const int NB = (N + 7) >> 3;
                const float4 q0 = reinterpret_cast<const float4*>(MG + L.f_bx0())[o];
                const float4* q1p = reinterpret_cast<const float4*>(MG + L.f_bx1()) + o * NB;
                mk = 0u;
#pragma unroll
                for (int j = 0; j < SPL; ++j) { my0[j] = 0u; my1[j] = 0u; }
#pragma unroll 1
                for (int b = 0; b < NB; ++b) {
                    const int ln = (8 * b) & 31;
                    const bool hi = SPL > 1 && 8 * b >= 32;
                    const float bl = __shfl_sync(FULL, hi ? rlx[SPL - 1] : rlx[0], ln), bh = __shfl_sync(FULL, hi ? rhx[SPL - 1] : rhx[0], ln);
                    const float cl = __shfl_sync(FULL, hi ? rly[SPL - 1] : rly[0], ln), ch = __shfl_sync(FULL, hi ? rhy[SPL - 1] : rhy[0], ln);
                    const float4 q1 = q1p[b];
                    const unsigned m0 = __ballot_sync(FULL, valid && !(bl > q0.y || bh < q0.x || cl > q0.w || ch < q0.z));
                    const unsigned m1 = __ballot_sync(FULL, valid && !(bl > q1.y || bh < q1.x || cl > q1.w || ch < q1.z));
#pragma unroll
                    for (int j = 0; j < SPL; ++j)
                        if ((kk[j] >> 3) == b) { my0[j] = m0; my1[j] = m1; }
                    mk |= m0 | m1;
                }
            } else {
                mk = __ballot_sync(FULL, it < L.Ndyn() && !(IM_E[it < L.Ndyn() ? it : 0] > dmax));
            }
            while (mk) {
                const int bit = __ffs(mk) - 1;
                const int i = base + bit;
                mk &= mk - 1;
                EllT a[SPL], b[SPL];
                double hl = 0.0;
                double hx[SPL], hy[SPL];      // team mode: a.hrx + b.hrx, a.hry + b.hry
                bool from_row = false;
                if constexpr (TEAM) {
                    // flagged ellipses arrive as rows (hinge, hrx, hry per step) from the workers
                    if (nflag < MPCB_TEAM_SLOTS) {
                        const double* row = team_ptrs(T, N).HROW + (size_t)nflag * N * 3;
#pragma unroll
                        for (int j = 0; j < SPL; ++j) {
                            hx[j] = 0.0; hy[j] = 0.0;
                            if (act[j]) {
                                const double* r = row + 3 * kk[j];
                                hl += r[0]; hx[j] = r[1]; hy[j] = r[2];
                            }
                        }
                        from_row = true;
                    }
                    ++nflag;
                }
                if (!from_row) {
#pragma unroll
                    for (int j = 0; j < SPL; ++j) {
                        const int k = act[j] ? kk[j] : N - 1;
                        a[j].hr = 0.0; b[j].hr = 0.0; a[j].hrx = 0.0; a[j].hry = 0.0; b[j].hrx = 0.0; b[j].hry = 0.0;
                        bool p0, p1;
                        if constexpr (BOX) { p0 = (my0[j] >> bit) & 1u; p1 = (my1[j] >> bit) & 1u; }
                        else { p0 = !(me0b[i * N + k] > Df[j]); p1 = !(metb[i * N + k] > Df[j]); }
                        if (p0) {                                  // t = 0 slot (one ellipse for all steps)
                            ellipse_terms(GRAD, e0 + i, L.Ndyn(), X[j], Y[j], a[j]);
                            if constexpr (!TEAM) {
                                cstj[j] += a[j].cost;
                                if (GRAD) { gx[j] += a[j].gx; gy[j] += a[j].gy; }
                            }
                        }
                        if (p1) {                                  // t = k+1 slot
                            ellipse_terms(GRAD, etb + k + i * N, L.Ndyn() * N, X[j], Y[j], b[j]);
                            if constexpr (!TEAM) {
                                cstj[j] += b[j].cost;
                                if (GRAD) { gx[j] += b[j].gx; gy[j] += b[j].gy; }
                            }
                        }
                        if (!act[j]) { a[j].hr = 0.0; b[j].hr = 0.0; }
                        hl += a[j].hr + b[j].hr;
                        if constexpr (TEAM) {
                            hx[j] = act[j] ? a[j].hrx + b[j].hrx : 0.0;
                            hy[j] = act[j] ? a[j].hry + b[j].hry : 0.0;
                        }
                    }
                }
                if (__any_sync(FULL, hl > 0.0)) {
                    // (team mode: with SP == +0.0 the catch-up adds exact zeros to non-negative sums)
                    if (TEAM && SP == 0.0 && !F2out) nxt = i;
#pragma unroll 1
                    for (; nxt < i; ++nxt) {                   // entries without a hinge equal SP
                        if (F2out && lane == 0) F2out[nxt] = SP;
                        f2sq = fma(SP, SP, f2sq);
                        sumF2 += SP;
                    }
                    const double F2i = SP + warp_sum(hl);
                    if (F2out && lane == 0) F2out[i] = F2i;
                    f2sq = fma(F2i, F2i, f2sq);
                    sumF2 += F2i;
                    nxt = i + 1;
                    anyF2 = true;
                    if (GRAD && F2i > 0.0) {
                        const double m = c * F2i;
#pragma unroll
                        for (int j = 0; j < SPL; ++j) {
                            if constexpr (TEAM) {
                                if (act[j]) { fx[j] = fma(m, hx[j], fx[j]); fy[j] = fma(m, hy[j], fy[j]); }
                            } else {
                                if (a[j].hr > 0.0) { fx[j] = fma(m, a[j].hrx, fx[j]); fy[j] = fma(m, a[j].hry, fy[j]); }
                                if (b[j].hr > 0.0) { fx[j] = fma(m, b[j].hrx, fx[j]); fy[j] = fma(m, b[j].hry, fy[j]); }
                            }
                        }
                    }
                }
            }
        }
        if (L.Ndyn() == 0) {
            f2sq = SP * SP;
            sumF2 = SP;
            if (F2out && lane == 0) F2out[0] = SP;
        } else if (anyF2) {
            if (TEAM && SP == 0.0 && !F2out) nxt = L.Ndyn();
#pragma unroll 1
            for (; nxt < L.Ndyn(); ++nxt) {
                if (F2out && lane == 0) F2out[nxt] = SP;
                f2sq = fma(SP, SP, f2sq);
                sumF2 += SP;
            }
        }
    }
#pragma unroll
    for (int j = 0; j < SPL; ++j) {
        if (!act[j]) { cstj[j] = 0.0; gx[j] = 0.0; gy[j] = 0.0; gvd[j] = 0.0; gwd[j] = 0.0; }
        cost += cstj[j];
    }

    EVAL_T(4);
    // ---- terminal cost on the last state (mpc_builder.py:148)
    double gthN = 0.0;
    {
        const double qN = q[5], qthN = q[6];
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

    // ---- F2's share of the gradient: the ellipse hinges collected above, then the polygon
    //      hinges (their sum enters every entry of F2)
    if (GRAD && anyF2) {
        const double m = c * sumF2;
#pragma unroll
        for (int j = 0; j < SPL; ++j) {
            gx[j] = fma(m, dSx[j], gx[j] + fx[j]);
            gy[j] = fma(m, dSy[j], gy[j] + fy[j]);
        }
    }

    // ---- accelerations: cost, ALM term on F1 (mpc_builder.py:156-169)
    double gFa[SPL], gFw[SPL];
    double dist2 = 0.0;
    {
        const double accp = q[8], waccp = q[9];
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
            const double za = acc + ya[j], zw = wacc + yw[j];
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
    // ONE reduction per evaluation: every lane folds its share of the ALM distance term into its
    // stage cost first, psi = sum_l (cost_l + c/2 dist2_l) + c/2 |F2|^2.  With c = 0 (the
    // evaluation the ALM step makes to read f(u), F1, F2) the sum is f itself, bit for bit;
    // a caller that wants f next to psi for c > 0 (the parity entry point K2) sets need_f and
    // pays the second reduction.
    const double hc = 0.5 * c;
    const double ps = warp_sum(fma(hc, dist2, cost));
    out.f = need_f ? warp_sum(cost) : ps;
    out.f2sq = f2sq;
    out.psi = fma(hc, f2sq, ps);
    EVAL_T(5);

    if (GRAD) {
        // adjoint: G = sum_{j>=k} g_j ; theta coupling via a second suffix sum
        double Gx[SPL], Gy[SPL], hh[SPL], Hex[SPL], Hin[SPL];
        // only the lane owning step N-1 holds a non-zero value: broadcasting it and adding +0.0
        // gives the bits the butterfly sum over {x, 0, ..., 0} would (x + 0.0 in every lane)
        const double gthN_all = __shfl_sync(FULL, gthN, (N - 1) & 31) + 0.0;
        suffix_incl<SPL>(gx, Gx, lane);
        suffix_incl<SPL>(gy, Gy, lane);
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
        EVAL_T(6);
    }
}

// Worker threads of a team (every warp of the CTA but the solver warps); t = thread index among the
// workers, `solvers` = the per-solver blocks (stride `sstride` doubles), NS of them take part.
// The pool serves one request at a time (worker thread 0 picks the next pending one, round robin):
// pass 1 (path segments, polygons and ellipses of the thread's step and group; group sums; ellipses
// with a raw hinge flagged), a barrier among the workers, pass 2 (totals over the groups; path
// minimum over the groups and its gradient; hinge rows of the flagged ellipses), a barrier, and the
// request is marked done.
template <int FIXED>
__device__ __forceinline__ void team_worker(const KParams& P, double* solvers, int sstride, int lb_doubles,
                                            int NS, TeamPool* pool, int t)
{
    const LayV<FIXED> L{&P.L};
    const int N = L.N(), Ndyn = L.Ndyn();
    const int G = FIXED != 0 ? team_groups(N, Ndyn) : P.team_G;   // (run-time dims: may be the forced latency mode)
    double* const PART = reinterpret_cast<double*>(pool) + 2;
    double* const PB = PART + 6 * G * N;
    int* const PI = reinterpret_cast<int*>(PB + G * N);
    const bool mine = t < G * N;
    const int g = t / N, k = mine ? t - g * N : 0;
    unsigned cls = 0u;                       // ellipses i with i % G == g (G divides 32)
    for (int b = g; b < 32; b += G) cls |= 1u << b;
    const bool CULL = P.cull != 0;
    const int NB = (N + 7) >> 3, blk = k >> 3;
    int last = NS - 1;
#ifdef MPCB_TEAM_PROF
    long long tp_idle = 0, tp_p1 = 0, tp_p2 = 0, tp_own = 0, tp_n = 0, tp_t = clock64();
#define TEAM_T(acc) do { const long long t_ = clock64(); acc += t_ - tp_t; tp_t = t_; } while (0)
#else
#define TEAM_T(acc) do { } while (0)
#endif
    for (;;) {
        if (t < 32) {
            // next pending request, round robin from the solver served last; -1 when every solver is
            // done.  The whole first worker warp polls (warp-uniform loop, broadcast loads).
            int pick = -2;
            while (pick == -2) {
                int nexit = 0;
                for (int q = 1; q <= NS; ++q) {
                    int s = last + q;
                    if (s >= NS) s -= NS;
                    TeamShared* Ts = reinterpret_cast<TeamShared*>(solvers + (size_t)s * sstride + lb_doubles);
                    if (Ts->req != Ts->done) { pick = s; break; }
                    nexit += Ts->exit_;
                }
                if (pick == -2) {
                    if (nexit == NS) pick = -1;
                    else { MPCB_SPIN_LEADER; }
                }
            }
            pick = __shfl_sync(FULL, pick, 0);     // one decision for the warp
            __threadfence_block();
            if (t == 0) pool->cur = pick;
        }
        __syncwarp();
        bar_sync(3, TEAM_WORKERS);
        TEAM_T(tp_idle);
        const int cur = pool->cur;
#ifdef MPCB_TEAM_PROF
        if (cur < 0 && (t & 31) == 0 && P.prof) {
            unsigned long long* o = P.prof + MPCB_WS_PROF_CTAS + (size_t)(blockIdx.x & (MPCB_WS_PROF_CTAS - 1)) * MPCB_WS_PROF_WARPS;
            o[NS + (t >> 5)] = (unsigned long long)tp_own;
            if (t == 0) { o[12] = (unsigned long long)tp_idle; o[13] = (unsigned long long)tp_p1; o[14] = (unsigned long long)tp_p2; o[15] = (unsigned long long)tp_n; }
        }
#endif
        if (cur < 0) return;
        last = cur;
        TeamShared* T = reinterpret_cast<TeamShared*>(solvers + (size_t)cur * sstride + lb_doubles);
        const TeamPtr tp = team_ptrs(T, N);
        const double* __restrict__ S = T->S;
        const bool GRAD = T->grad != 0;
        const float* MG = reinterpret_cast<const float*>(S + L.o_mg());
        const double* e0 = S + L.o_e0();
        const double* etb = S + L.o_et() + k;
        const float4* bx0 = reinterpret_cast<const float4*>(MG + L.f_bx0());
        const float4* bx1 = reinterpret_cast<const float4*>(MG + L.f_bx1()) + blk;
        const double x = tp.X[k], y = tp.Y[k];
        // bounding box of the robot's positions over this step's block of 8 steps, rounded outwards
        float rx0, rx1, ry0, ry1;
        {
            const int ka = blk << 3, kb = ka + 8 < N ? ka + 8 : N;
            double a0 = tp.X[ka], a1 = a0, b0 = tp.Y[ka], b1 = b0;
            bool bad = !(a0 == a0) || !(b0 == b0);
#pragma unroll 1
            for (int q = ka + 1; q < kb; ++q) {
                const double xq = tp.X[q], yq = tp.Y[q];
                bad |= !(xq == xq) || !(yq == yq);
                a0 = xq < a0 ? xq : a0; a1 = xq > a1 ? xq : a1;
                b0 = yq < b0 ? yq : b0; b1 = yq > b1 ? yq : b1;
            }
            rx0 = __double2float_rd(a0); rx1 = __double2float_ru(a1);
            ry0 = __double2float_rd(b0); ry1 = __double2float_ru(b1);
            if (bad || !CULL) { rx0 = ry0 = -__int_as_float(0x7f800000); rx1 = ry1 = __int_as_float(0x7f800000); }
        }
        if (mine) {
            // ---- pass 1a: reference path, segments k+g, k+g+G, ... (mpc_cost.py:84-95)
            {
                const double* sg = S + L.o_seg();
                double best = INFINITY;
                int ib = N;
#pragma unroll 1
                for (int i = k + g; i < N; i += G) {
                    const double ex = x - sg[i], ey = y - sg[N + i];
                    const double ddx = sg[2 * N + i], ddy = sg[3 * N + i];
                    const double th_ = fma(ex, ddx, ey * ddy) * sg[4 * N + i];
                    const double ts_ = clamp01(th_);
                    const double vx = fma(ts_, ddx, -ex), vy = fma(ts_, ddy, -ey);
                    const double d2 = fma(vx, vx, vy * vy);
                    if (d2 < best) { best = d2; ib = i; }
                }
                PB[g * N + k] = best;
                PI[g * N + k] = ib;
            }
            // ---- pass 1b: polygons i % G == g (mpc_builder.py:100-108), then ellipses i % G == g
            double pc = 0.0, pgx = 0.0, pgy = 0.0, psp = 0.0, pspx = 0.0, pspy = 0.0;
            {
                const double qs = S[L.o_qstc() + k];
                const double* pe = S + L.o_poly();
#pragma unroll 1
                for (int i = g; i < L.Nstc(); i += G) {
                    double dIx, dIy;
                    const double I = polygon_ind(GRAD, pe + i * 3 * L.nedge(), L.nedge(), x, y, dIx, dIy);
                    if (I > 0.0) {
                        pc = fma(qs, I * I, pc);
                        psp += I;
                        if (GRAD) {
                            const double m = 2.0 * qs * I;
                            pgx = fma(m, dIx, pgx);
                            pgy = fma(m, dIy, pgy);
                            pspx += dIx;
                            pspy += dIy;
                        }
                    }
                }
            }
            // Two phases per chunk of 32 of the thread's ellipses: first every bounding-box test (a short
            // loop whose loads are all in flight together: the boxes come from L2 as often as from L1),
            // then the survivors in index order - the same terms in the same order as one fused loop.
#pragma unroll 1
            for (int i0 = g; i0 < Ndyn; i0 += 32 * G) {
                // an ellipse whose bounding box misses the block's robot box is exactly zero at
                // every step of the block (NaN boxes compare false: never skipped)
                unsigned m0 = 0u, m1 = 0u;
#pragma unroll 4
                for (int j = 0; j < 32; ++j) {
                    const int i = i0 + j * G;
                    if (i < Ndyn) {
                        const float4 q0 = bx0[i], q1 = bx1[i * NB];
                        if (!(rx0 > q0.y || rx1 < q0.x || ry0 > q0.w || ry1 < q0.z)) m0 |= 1u << j;
                        if (!(rx0 > q1.y || rx1 < q1.x || ry0 > q1.w || ry1 < q1.z)) m1 |= 1u << j;
                    }
                }
                unsigned m = m0 | m1;
                while (m) {
                    const int j = __ffs(m) - 1;
                    m &= m - 1;
                    const int i = i0 + j * G;
                    EllT a, b;
                    a.hr = 0.0; b.hr = 0.0;
                    if ((m0 >> j) & 1u) {                    // t = 0 slot
                        ellipse_terms(GRAD, e0 + i, Ndyn, x, y, a);
                        pc += a.cost;
                        if (GRAD) { pgx += a.gx; pgy += a.gy; }
                    }
                    if ((m1 >> j) & 1u) {                    // t = k+1 slot
                        ellipse_terms(GRAD, etb + i * N, Ndyn * N, x, y, b);
                        pc += b.cost;
                        if (GRAD) { pgx += b.gx; pgy += b.gy; }
                    }
                    if (a.hr > 0.0 || b.hr > 0.0) atomicOr(&T->hit[i >> 5], 1u << (i & 31));
                }
            }
            double* q = PART + g * N + k;
            q[0] = pc; q[G * N] = pgx; q[2 * G * N] = pgy;
            q[3 * G * N] = psp; q[4 * G * N] = pspx; q[5 * G * N] = pspy;
        }
        __syncwarp();
#ifdef MPCB_TEAM_PROF
        tp_own += clock64() - tp_t;
#endif
        bar_sync(3, TEAM_WORKERS);
        TEAM_T(tp_p1);
        // ---- pass 2a: totals over the groups, in group order from +0.0 (6 quantities x N steps)
#pragma unroll 1
        for (int idx = t; idx < 6 * N; idx += TEAM_WORKERS) {
            const int qn = idx / N, kq = idx - qn * N;
            const double* q = PART + qn * G * N + kq;
            double tot = 0.0;
#pragma unroll 1
            for (int gg = 0; gg < G; ++gg) tot += q[gg * N];
            tp.SUM[idx] = tot;
        }
        if (mine) {
            // ---- pass 2b (group G-1: the threads pass 2a leaves idle first): path minimum over the
            //      groups (ties: lowest index), cost, gradient
            if (g == G - 1) {
                const double* sg = S + L.o_seg();
                const double qrpd = S[L.o_hdr() + H_Q + 7];
                double best = INFINITY;
                int ib = k;
#pragma unroll 1
                for (int gg = 0; gg < G; ++gg) {
                    const double bq = PB[gg * N + k];
                    const int iq = PI[gg * N + k];
                    if (bq < best || (bq == best && iq < ib)) { best = bq; ib = iq; }
                }
                double ggx = 0.0, ggy = 0.0;
                if (GRAD) {
                    const double ex = x - sg[ib], ey = y - sg[N + ib];
                    const double ddx = sg[2 * N + ib], ddy = sg[3 * N + ib], inv = sg[4 * N + ib];
                    const double th_ = fma(ex, ddx, ey * ddy) * inv;
                    const double ts_ = clamp01(th_);
                    const double vx = fma(ts_, ddx, -ex), vy = fma(ts_, ddy, -ey);
                    const double dt = (th_ > 0.0 && th_ < 1.0) ? 1.0 : ((th_ == 0.0 || th_ == 1.0) ? 0.5 : 0.0);
                    const double vd = fma(vx, ddx, vy * ddy) * dt * inv;
                    ggx = 2.0 * qrpd * fma(vd, ddx, -vx);
                    ggy = 2.0 * qrpd * fma(vd, ddy, -vy);
                }
                tp.PATH[k] = best * qrpd; tp.PATH[N + k] = ggx; tp.PATH[2 * N + k] = ggy;
            }
            // ---- pass 2c: hinge rows of the flagged ellipses of this group (row = rank of the ellipse
            //      among the flagged ones), as far as there are rows
            int below = 0;
#pragma unroll 1
            for (int base = 0; base < Ndyn; base += 32) {
                const unsigned hw = T->hit[base >> 5];
                unsigned m = hw & cls;
#pragma unroll 1
                while (m) {
                    const int bit = __ffs(m) - 1;
                    const int i = base + bit;
                    m &= m - 1;
                    const int slot = below + __popc(hw & ((1u << bit) - 1u));
                    if (slot >= MPCB_TEAM_SLOTS) break;
                    const float4 q0 = bx0[i], q1 = bx1[i * NB];
                    const bool in0 = !(rx0 > q0.y || rx1 < q0.x || ry0 > q0.w || ry1 < q0.z);
                    const bool in1 = !(rx0 > q1.y || rx1 < q1.x || ry0 > q1.w || ry1 < q1.z);
                    EllT a, b;
                    a.hr = 0.0; a.hrx = 0.0; a.hry = 0.0; b.hr = 0.0; b.hrx = 0.0; b.hry = 0.0;
                    if (in0) ellipse_terms(GRAD, e0 + i, Ndyn, x, y, a);
                    if (in1) ellipse_terms(GRAD, etb + i * N, Ndyn * N, x, y, b);
                    double* r = tp.HROW + ((size_t)slot * N + k) * 3;
                    r[0] = a.hr + b.hr; r[1] = a.hrx + b.hrx; r[2] = a.hry + b.hry;
                }
                below += __popc(hw);
            }
        }
        __syncwarp();
        bar_sync(3, TEAM_WORKERS);
        TEAM_T(tp_p2);
#ifdef MPCB_TEAM_PROF
        ++tp_n;
#endif
        if (t == 0) {
            __threadfence_block();
            T->done = T->req;
        }
#ifdef MPCB_TEAM_BAR_WAIT
        if (t < 32) {
            __syncwarp();
            asm volatile("bar.arrive %0, %1;" ::"r"(8 + cur), "n"(64) : "memory");
        }
#endif
    }
}

}  // namespace mpcb
