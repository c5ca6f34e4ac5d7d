// mpcb.cu — kernels and the C-ABI (include/mpcb.h) of the batched NMPC solver.
//
//   K3 stage_kernel : raw parameter rows p (reference AoS layout,
//                     mpc_builder.py:47-60) -> structure-of-arrays scenario
//                     blocks in the workspace (cos/sin of ellipse angles,
//                     inverse squared radii, weights folded in: done once per
//                     scenario instead of once per cost evaluation).
//   K2 eval_kernel  : psi / grad psi / F1 / F2 for B instances (parity tests).
//   K1 solve_kernel_queue (default): persistent CTAs, one per SM, one work queue per CTA
//                     (queue q owns scenarios q, q+grid, ...; idle warps steal); every warp
//                     runs the whole ALM/PANOC solve of one instance out of registers and its
//                     shared-memory scratch, reading the scenario block through L1.
//      solve_kernel (MPCB_MODE=1, kept for comparison): CTA-synchronous groups that
//                     TMA-bulk-load their scenario block(s) into shared memory.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -fmad=false
// (-fmad=false: only the explicit fma() calls fuse, so the arithmetic does not
// depend on the compiler's contraction choices).
#include <cuda_runtime.h>
#include <stdio.h>
#include <string.h>
#include <stdlib.h>

#include "mpcb_device.cuh"
#include "mpcb_solver.cuh"
#include "mpcb_sim.cuh"

#ifndef MPCB_QTHREADS
#define MPCB_QTHREADS 256
#endif
#ifndef MPCB_QTHREADS_FIXED
#define MPCB_QTHREADS_FIXED 512   // 16 warps x 128 registers (solver state parked in shared memory during the
                                  // evaluation, per-CTA queues keep L1 on 2-3 scenario blocks): best measured
#endif
#ifndef MPCB_QTHREADS_FIXED2
#define MPCB_QTHREADS_FIXED2 384  // 40 ellipses: 148 registers without spills, 12 warps
#endif
#ifndef MPCB_MIN_CTAS
#define MPCB_MIN_CTAS 1
#endif
// threads per CTA of the queue kernel for a compiled-in dimension set (0: run-time dims)
constexpr int qthreads(int fixed)
{
    return fixed == 1 ? MPCB_QTHREADS_FIXED : fixed == 2 ? MPCB_QTHREADS_FIXED2 : MPCB_QTHREADS;
}
using namespace mpcb;

namespace {

thread_local char g_err[256] = "";

#define CUDA_TRY(x)                                                                  \
    do {                                                                             \
        cudaError_t e_ = (x);                                                        \
        if (e_ != cudaSuccess) {                                                     \
            snprintf(g_err, sizeof(g_err), "%s: %s", #x, cudaGetErrorString(e_));    \
            return MPCB_E_CUDA;                                                      \
        }                                                                            \
    } while (0)

// ----------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p)
{
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// TMA 1-D bulk copy global -> shared, completion on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void tma_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar)
{
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(dst)),
        "l"(src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

// ------------------------------------------------------------------ K3: staging
__device__ __forceinline__ double hypot_plain(double a, double b) { return sqrt(a * a + b * b); }
// conservative float margin: dist - R, shrunk by 1e-9 relative + absolute, rounded down
__device__ __forceinline__ float margin_f(double dist, double R)
{
    const double m = dist - R * (1.0 + 1e-9) - 1e-9;
    if (m == INFINITY) return __int_as_float(0x7f800000);   // "never active" (zero-row polygon): INF - INF*1e-9 is NaN
    return __double2float_rd(m - fabs(m) * 1e-9);
}

__global__ void __launch_bounds__(128) stage_kernel(const KParams P, const double* __restrict__ p,
                                                    double* __restrict__ staged, float* __restrict__ keys)
{
    const Lay& L = P.L;
    const int N = L.N;
    for (int s = blockIdx.x; s < P.n_p; s += gridDim.x) {
        const double* pr = p + (size_t)s * L.np;
        double* S = staged + (size_t)s * L.total;
        float* MG = reinterpret_cast<float*>(S + L.o_mg);
        const int tid = threadIdx.x, nt = blockDim.x;
        // header
        for (int i = tid; i < H_SIZE; i += nt) {
            double v = 0.0;
            if (i <= H_S0T) v = pr[L.p_s0 + i];
            else if (i <= H_UM1W) v = pr[L.p_um1 + (i - H_UM1V)];
            else if (i <= H_SNT) v = pr[L.p_sN + (i - H_SNX)];
            else if (i < H_Q + 10) v = pr[L.p_q + (i - H_Q)];
            S[L.o_hdr + i] = v;
        }
        for (int k = tid; k < N; k += nt) {
            S[L.o_rv + k] = pr[L.p_rv + k];
            S[L.o_qstc + k] = pr[L.p_qstc + k];
            // reference polyline: rows 0..N-1 of r_s, row N duplicates row N-1
            // (mpc_builder.py:68-69); segment k joins rows k and k+1
            const int k2 = k + 1 < N ? k + 1 : N - 1;
            const double ax = pr[L.p_rs + 3 * k], ay = pr[L.p_rs + 3 * k + 1];
            const double dx = pr[L.p_rs + 3 * k2] - ax, dy = pr[L.p_rs + 3 * k2 + 1] - ay;
            S[L.o_seg + k] = ax;
            S[L.o_seg + N + k] = ay;
            S[L.o_seg + 2 * N + k] = dx;
            S[L.o_seg + 3 * N + k] = dy;
            S[L.o_seg + 4 * N + k] = 1.0 / (dx * dx + dy * dy + 1e-16);
            // unit tangent, shortened so that |t| <= 1 survives rounding (zero for a padded segment)
            const double dl = sqrt(dx * dx + dy * dy);
            S[L.o_seg + 5 * N + k] = dl > 1e-9 ? dx / dl * (1.0 - 1e-9) : 0.0;
            S[L.o_seg + 6 * N + k] = dl > 1e-9 ? dy / dl * (1.0 - 1e-9) : 0.0;
        }
        for (int r = tid; r < L.Nother; r += nt) {
            S[L.o_c0 + r] = pr[L.p_c0 + 3 * r];
            S[L.o_c0 + L.Nother + r] = pr[L.p_c0 + 3 * r + 1];
        }
        for (int i = tid; i < L.Nother * N; i += nt) {   // i = r*N + k; c is robot-major
            const int r = i / N, k = i - r * N;
            S[L.o_c + i] = pr[L.p_c + r * 3 * N + 3 * k];
            S[L.o_c + L.Nother * N + i] = pr[L.p_c + r * 3 * N + 3 * k + 1];
        }
        for (int i = tid; i < L.Nstc * L.nedge; i += nt) {
            const int pl = i / L.nedge, e = i - pl * L.nedge;
            const double* q = pr + L.p_os + pl * 3 * L.nedge;
            S[L.o_poly + 3 * i] = q[e];
            S[L.o_poly + 3 * i + 1] = -q[L.nedge + e];
            S[L.o_poly + 3 * i + 2] = -q[2 * L.nedge + e];
        }
        for (int i = tid; i < L.Ndyn * (N + 1); i += nt) {   // i = obstacle*(N+1) + t
            const int ob = i / (N + 1), t = i - ob * (N + 1);
            const double* q = pr + L.p_od + (size_t)i * 6;
            const double rx = q[2], ry = q[3];
            const double rxi = t == 0 ? rx + P.vmargin + P.smargin : rx + P.vmargin;
            const double ryi = t == 0 ? ry + P.vmargin + P.smargin : ry + P.vmargin;
            const double wgt = t == 0 ? 1000.0 : pr[L.p_qdyn + t - 1];
            double sa, ca;
            sincos_cw(q[4], &sa, &ca);
            double f[EF];
            f[E_CX] = q[0]; f[E_CY] = q[1]; f[E_CA] = ca; f[E_SA] = sa;
            f[E_I1I] = 1.0 / ((rxi + 1e-6) * (rxi + 1e-6));
            f[E_I2I] = 1.0 / ((ryi + 1e-6) * (ryi + 1e-6));
            f[E_I1R] = 1.0 / ((rx + 1e-6) * (rx + 1e-6));
            f[E_I2R] = 1.0 / ((ry + 1e-6) * (ry + 1e-6));
            f[E_WAL] = wgt * q[5];
#pragma unroll
            for (int m = 0; m < EF; ++m) {
                if (t == 0) S[L.o_e0 + m * L.Ndyn + ob] = f[m];
                else S[L.o_et + (m * L.Ndyn + ob) * N + (t - 1)] = f[m];
            }
        }
        __syncthreads();   // the anchors (segment start points) and polygon rows are staged
        // ---- culling tables, all relative to the anchors A_k = r_s[k]
        const double* sg = S + L.o_seg;
        for (int i = tid; i < L.Ndyn * N; i += nt) {          // i = ob*N + k
            const int ob = i / N, k = i - ob * N;
            const double ax = sg[k], ay = sg[N + k];
            {   // t = 0 slot
                const double* q = pr + L.p_od + (size_t)(ob * (N + 1)) * 6;
                const double Rb = fmax(fabs(q[2] + P.vmargin + P.smargin + 1e-6), fabs(q[3] + P.vmargin + P.smargin + 1e-6));
                MG[L.f_e0 + i] = margin_f(hypot_plain(q[0] - ax, q[1] - ay), Rb);
            }
            {   // t = k+1 slot
                const double* q = pr + L.p_od + (size_t)(ob * (N + 1) + k + 1) * 6;
                const double Rb = fmax(fabs(q[2] + P.vmargin + 1e-6), fabs(q[3] + P.vmargin + 1e-6));
                MG[L.f_et + i] = margin_f(hypot_plain(q[0] - ax, q[1] - ay), Rb);
            }
        }
        // bounding boxes of the inflated ellipses, per obstacle (t = 0 slot) and per obstacle and
        // block of 8 steps (t = k+1 slots): the position-based culling test of the team kernels
        // (a robot that lags behind its reference makes every anchor-based margin useless)
        {
            const int NB = (N + 7) / 8;
            for (int idx = tid; idx < L.Ndyn * (NB + 1); idx += nt) {
                const int ob = idx / (NB + 1), b = idx - ob * (NB + 1) - 1;     // b = -1: the t = 0 slot
                const int k0 = b < 0 ? -1 : 8 * b, k1 = b < 0 ? 0 : (8 * b + 8 < N ? 8 * b + 8 : N);
                float x0 = __int_as_float(0x7f800000), x1 = -x0, y0 = x0, y1 = -x0;
                bool bad = false;
                for (int k = k0; k < k1; ++k) {
                    const double* q = pr + L.p_od + (size_t)(ob * (N + 1) + k + 1) * 6;
                    const double infl = b < 0 ? P.vmargin + P.smargin : P.vmargin;
                    const double Rb = fmax(fabs(q[2] + infl + 1e-6), fabs(q[3] + infl + 1e-6)) * (1.0 + 1e-9) + 1e-9;
                    const double lx = q[0] - Rb, hx = q[0] + Rb, ly = q[1] - Rb, hy = q[1] + Rb;
                    bad |= !(lx == lx) || !(hx == hx) || !(ly == ly) || !(hy == hy);
                    x0 = fminf(x0, __double2float_rd(lx - fabs(lx) * 1e-9)); x1 = fmaxf(x1, __double2float_ru(hx + fabs(hx) * 1e-9));
                    y0 = fminf(y0, __double2float_rd(ly - fabs(ly) * 1e-9)); y1 = fmaxf(y1, __double2float_ru(hy + fabs(hy) * 1e-9));
                }
                if (bad) { x0 = x1 = y0 = y1 = __int_as_float(0x7fc00000); }   // NaN: never culled
                float* o = b < 0 ? MG + L.f_bx0 + 4 * ob : MG + L.f_bx1 + 4 * (ob * NB + b);
                o[0] = x0; o[1] = x1; o[2] = y0; o[3] = y1;
            }
        }
        for (int i = tid; i < L.Nother * N; i += nt) {         // i = r*N + k
            const int r = i / N, k = i - r * N;
            const double ax = sg[k], ay = sg[N + k];
            MG[L.f_c0 + i] = margin_f(hypot_plain(pr[L.p_c0 + 3 * r] - ax, pr[L.p_c0 + 3 * r + 1] - ay), P.dsafe);
            MG[L.f_c + i] = margin_f(hypot_plain(pr[L.p_c + r * 3 * N + 3 * k] - ax,
                                                 pr[L.p_c + r * 3 * N + 3 * k + 1] - ay), P.dsafe);
        }
        for (int i = tid; i < L.Nstc * N; i += nt) {           // i = pl*N + k
            // a polygon is exactly zero wherever one half-plane value is <= 0: the bound is the
            // largest distance by which the anchor violates an edge (zero rows: never active)
            const int pl = i / N, k = i - pl * N;
            const double ax = sg[k], ay = sg[N + k];
            const double* q = pr + L.p_os + pl * 3 * L.nedge;
            double m = -INFINITY;
            for (int e = 0; e < L.nedge; ++e) {
                const double a0 = q[L.nedge + e], a1 = q[2 * L.nedge + e];
                const double r0 = q[e] - a0 * ax - a1 * ay;
                const double an = sqrt(a0 * a0 + a1 * a1);
                if (an > 0.0) m = fmax(m, -r0 / an);
                else if (r0 <= 0.0) m = INFINITY;
            }
            MG[L.f_poly + i] = margin_f(m, 0.0);
        }
        // bounding box of every polygon (position-based culling: the indicator is a product of
        // max(0, half-plane value), exactly zero outside the polygon).  Vertices = feasible pairwise
        // intersections of the edge lines; an unbounded polygon (normals that do not positively span
        // the plane) gets an infinite box, a zero-row slot an empty one.
        for (int pl = tid; pl < L.Nstc; pl += nt) {
            const double* q = pr + L.p_os + pl * 3 * L.nedge;
            const float PINF = __int_as_float(0x7f800000);
            float bx[4] = {PINF, -PINF, PINF, -PINF};           // empty
            bool zero_row = false, bad = false;
            double ang[MPCB_MAX_EDGE];
            bool real[MPCB_MAX_EDGE];          // a genuine half-plane (a constant row b > 0 is only a factor)
            int nreal = 0;
            for (int e = 0; e < L.nedge; ++e) {
                const double a0 = q[L.nedge + e], a1 = q[2 * L.nedge + e], b = q[e];
                bad |= !(a0 == a0) || !(a1 == a1) || !(b == b);
                real[e] = !(a0 == 0.0 && a1 == 0.0);
                if (!real[e] && b <= 0.0) zero_row = true;                 // this edge is never positive
                nreal += real[e];
                ang[e] = atan2(a1, a0);
            }
            bool bounded = !bad && nreal >= 3;
            if (bounded && !zero_row) {
                // largest angular gap between consecutive outward normals must be < pi
                for (int e = 0; e < L.nedge && bounded; ++e) {
                    if (!real[e]) continue;
                    double gap = 7.0;
                    for (int f = 0; f < L.nedge; ++f) {
                        if (f == e || !real[f]) continue;
                        double d = ang[f] - ang[e];
                        while (d <= 0.0) d += 6.283185307179586;
                        gap = d < gap ? d : gap;
                    }
                    if (gap >= 3.141592653589793 - 1e-9) bounded = false;
                }
                if (bounded) {
                    int nv = 0;
                    for (int e = 0; e < L.nedge; ++e)
                        for (int f = e + 1; f < L.nedge; ++f) {
                            const double a0 = q[L.nedge + e], a1 = q[2 * L.nedge + e], b = q[e];
                            const double c0 = q[L.nedge + f], c1 = q[2 * L.nedge + f], d = q[f];
                            const double det = a0 * c1 - a1 * c0;
                            if (fabs(det) < 1e-14 * (fabs(a0 * c1) + fabs(a1 * c0)) || det == 0.0) continue;
                            const double vx = (b * c1 - a1 * d) / det, vy = (a0 * d - b * c0) / det;
                            bool feas = true;
                            for (int g = 0; g < L.nedge; ++g) {
                                const double r = q[g] - q[L.nedge + g] * vx - q[2 * L.nedge + g] * vy;
                                const double sc_ = fabs(q[g]) + fabs(q[L.nedge + g] * vx) + fabs(q[2 * L.nedge + g] * vy);
                                if (r < -1e-9 * sc_ - 1e-12) feas = false;
                            }
                            if (!feas) continue;
                            ++nv;
                            const double mx = 1e-6 + 1e-9 * fabs(vx), my = 1e-6 + 1e-9 * fabs(vy);
                            bx[0] = fminf(bx[0], __double2float_rd(vx - mx)); bx[1] = fmaxf(bx[1], __double2float_ru(vx + mx));
                            bx[2] = fminf(bx[2], __double2float_rd(vy - my)); bx[3] = fmaxf(bx[3], __double2float_ru(vy + my));
                        }
                    if (nv < 3) bounded = false;
                }
            }
            if (!zero_row && !bounded) { bx[0] = -PINF; bx[1] = PINF; bx[2] = -PINF; bx[3] = PINF; }   // never culled
            float* o = MG + L.f_pbx + 4 * pl;
            o[0] = bx[0]; o[1] = bx[1]; o[2] = bx[2]; o[3] = bx[3];
        }
        // reference path: (k, i) -> min over segments i' >= i of dist(A_k, segment i')
        for (int k = tid; k < N; k += nt) {
            const double ax = sg[k], ay = sg[N + k];
            double run = INFINITY;
            for (int i = N - 1; i >= 0; --i) {
                const double ex = ax - sg[i], ey = ay - sg[N + i];
                const double ddx = sg[2 * N + i], ddy = sg[3 * N + i];
                double t = (ex * ddx + ey * ddy) * sg[4 * N + i];
                t = fmin(fmax(t, 0.0), 1.0);
                const double vx = t * ddx - ex, vy = t * ddy - ey;
                run = fmin(run, sqrt(vx * vx + vy * vy));
                MG[L.f_seg + k * N + i] = margin_f(run, 0.0);
            }
            // directional twin: the projection on t_k is linear along a segment, so its minimum
            // over a segment sits at an end point
            const double tx = sg[5 * N + k], ty = sg[6 * N + k];
            double run2 = INFINITY;
            for (int i = N - 1; i >= 0; --i) {
                const double ex = sg[i] - ax, ey = sg[N + i] - ay;
                const double p0 = ex * tx + ey * ty;
                const double p1 = (ex + sg[2 * N + i]) * tx + (ey + sg[3 * N + i]) * ty;
                run2 = fmin(run2, fmin(p0, p1));
                MG[L.f_seg2 + k * N + i] = margin_f(run2, 0.0);
            }
        }
        __syncthreads();
        for (int i = tid; i < L.Ndyn + L.Nstc + 2 * L.Nother; i += nt) {   // per-item minimum over the steps
            int it = i, base, base2 = -1;
            if (it < L.Ndyn) { base = L.f_e0; base2 = L.f_et; }
            else if ((it -= L.Ndyn) < L.Nstc) base = L.f_poly;
            else if ((it -= L.Nstc) < L.Nother) base = L.f_c0;
            else { it -= L.Nother; base = L.f_c; }
            float m = __int_as_float(0x7f800000);
            for (int k = 0; k < N; ++k) {
                float t = MG[base + it * N + k];
                m = (t < m || t != t) ? t : m;                 // NaN poisons the minimum: never skipped
                if (base2 >= 0) {
                    t = MG[base2 + it * N + k];
                    m = (t < m || t != t) ? t : m;
                }
            }
            MG[L.f_imin + i] = m;
        }
        __syncthreads();
        // difficulty key of the scenario: how close an (active) ellipse comes to the reference path.
        // Scenarios whose path is blocked are the ones that use up the iteration caps; the solve
        // kernel takes the scenarios in the order of this key, so the long solves start first and
        // the launch does not end with a few warps grinding through them (order_kernel).
        if (keys && tid == 0) {
            float key = __int_as_float(0x7f800000);
            for (int i = 0; i < L.Ndyn; ++i) {
                if (S[L.o_e0 + E_WAL * L.Ndyn + i] == 0.0) continue;      // unused slot (alpha = 0)
                const float m = MG[L.f_imin + i];
                key = (m < key || m != m) ? m : key;
            }
            keys[s] = key;
        }
        __syncthreads();
    }
}

// Counting sort of the scenarios by difficulty key (one CTA; 128 bins of 1/16 m; NaN first): order[r]
// = the scenario the r-th queue slot works on.  The order inside a bin is not deterministic, which
// only moves WHEN an instance is solved, never what comes out.
constexpr int ORDER_BINS = 128;
__global__ void __launch_bounds__(1024) order_kernel(int n_p, const float* __restrict__ keys, int* __restrict__ order)
{
    __shared__ int hist[ORDER_BINS];
    auto bin = [](float k) {
        if (!(k == k)) return 0;
        const float b = floorf((k + 1.0f) * 16.0f);
        return b < 0.f ? 0 : (b > (float)(ORDER_BINS - 1) ? ORDER_BINS - 1 : (int)b);
    };
    for (int i = threadIdx.x; i < ORDER_BINS; i += blockDim.x) hist[i] = 0;
    __syncthreads();
    for (int s = threadIdx.x; s < n_p; s += blockDim.x) atomicAdd(&hist[bin(keys[s])], 1);
    __syncthreads();
    if (threadIdx.x == 0) {
        int run = 0;
        for (int i = 0; i < ORDER_BINS; ++i) { const int c = hist[i]; hist[i] = run; run += c; }
    }
    __syncthreads();
    for (int s = threadIdx.x; s < n_p; s += blockDim.x) order[atomicAdd(&hist[bin(keys[s])], 1)] = s;
}

// CTA-shared bookkeeping at the front of dynamic shared memory
struct CtaShared {
    uint64_t bar;
    int work;
    int pad;
};

// load the scenario blocks [sc0, sc1] into shared memory with TMA bulk copies
__device__ __forceinline__ void load_scenarios(const KParams& P, const double* staged, double* dst,
                                               int sc0, int sc1, uint64_t* bar, uint32_t& phase)
{
    const uint32_t blk = (uint32_t)P.L.total * 8u;
    if (threadIdx.x == 0) {
        mbar_expect_tx(bar, blk * (uint32_t)(sc1 - sc0 + 1));
        for (int s = sc0; s <= sc1; ++s) {
            const char* src = reinterpret_cast<const char*>(staged + (size_t)s * P.L.total);
            char* d = reinterpret_cast<char*>(dst + (size_t)(s - sc0) * P.L.total);
            for (uint32_t off = 0; off < blk; off += 32768u) {
                const uint32_t n = blk - off < 32768u ? blk - off : 32768u;
                tma_bulk_g2s(d + off, src + off, n, bar);
            }
        }
    }
    mbar_wait(bar, phase);
    phase ^= 1u;
}

// ---------------------------------------------------------------- K2: evaluation
template <int SPL, bool SMEM>
__global__ void __launch_bounds__(256) eval_kernel(const KParams P, const double* __restrict__ staged,
                                                   const double* __restrict__ u,
                                                   const double* __restrict__ y,
                                                   const double* __restrict__ c, double* f,
                                                   double* psi, double* grad, double* F1, double* F2)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    CtaShared* cs = reinterpret_cast<CtaShared*>(smem_raw);
    double* scn = reinterpret_cast<double*>(smem_raw + 16);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int N = P.L.N;
    const int b0 = blockIdx.x * P.warps;
    const int blast = min(b0 + P.warps, P.B) - 1;
    const int sc0 = b0 / P.starts, sc1 = blast / P.starts;
    uint32_t phase = 0;
    if (SMEM) {
        if (threadIdx.x == 0) {
            mbar_init(&cs->bar, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
        load_scenarios(P, staged, scn, sc0, sc1, &cs->bar, phase);
    }
    const int b = b0 + warp;
    if (b >= P.B) return;
    const int sc = b / P.starts;
    const double* S = SMEM ? scn + (size_t)(sc - sc0) * P.L.total : staged + (size_t)sc * P.L.total;

    double v[SPL], w[SPL], ya[SPL], yw[SPL];
    bool act[SPL];
#pragma unroll
    for (int j = 0; j < SPL; ++j) {
        const int k = lane + 32 * j;
        act[j] = k < N;
        v[j] = 0.0; w[j] = 0.0; ya[j] = 0.0; yw[j] = 0.0;
        if (act[j]) {
            v[j] = u[(size_t)b * 2 * N + 2 * k];
            w[j] = u[(size_t)b * 2 * N + 2 * k + 1];
            if (y) { ya[j] = y[(size_t)b * 2 * N + k]; yw[j] = y[(size_t)b * 2 * N + N + k]; }
        }
    }
    const double cc = c ? c[b] : P.c_init;
#pragma unroll
    for (int j = 0; j < SPL; ++j) { ya[j] = ya[j] / fmax(cc, 1.0); yw[j] = yw[j] / fmax(cc, 1.0); }
    const int n2 = P.L.Ndyn > 0 ? P.L.Ndyn : 1;
    EvalOut<SPL> o;
    eval_psi<SPL, 0>(P, S, v, w, cc, ya, yw, true, o, lane, F2 ? F2 + (size_t)b * n2 : nullptr, true);
    if (lane == 0) {
        if (f) f[b] = o.f;
        if (psi) psi[b] = o.psi;
    }
    // F1 = [acc; wacc]  (mpc_builder.py:158-160)
    double vc = S[P.L.o_hdr + H_UM1V], wc = S[P.L.o_hdr + H_UM1W];
#pragma unroll
    for (int j = 0; j < SPL; ++j) {
        const int k = lane + 32 * j;
        double vp = __shfl_up_sync(FULL, v[j], 1), wp = __shfl_up_sync(FULL, w[j], 1);
        if (lane == 0) { vp = vc; wp = wc; }
        if (SPL > 1) { vc = __shfl_sync(FULL, v[j], 31); wc = __shfl_sync(FULL, w[j], 31); }
        if (act[j]) {
            if (grad) {
                grad[(size_t)b * 2 * N + 2 * k] = o.gv[j];
                grad[(size_t)b * 2 * N + 2 * k + 1] = o.gw[j];
            }
            if (F1) {
                F1[(size_t)b * 2 * N + k] = (v[j] - vp) * P.inv_ts;
                F1[(size_t)b * 2 * N + N + k] = (w[j] - wp) * P.inv_ts;
            }
        }
    }
}

// ------------------------------------------------------------------- K1: solve
template <int SPL, bool SMEM>
__global__ void __launch_bounds__(256) solve_kernel(const KParams P, const double* __restrict__ staged,
                                                    const SolveIO io, int* __restrict__ counter)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    CtaShared* cs = reinterpret_cast<CtaShared*>(smem_raw);
    double* scn = reinterpret_cast<double*>(smem_raw + 16);
    double* lb_all = scn + (SMEM ? (size_t)P.nsc * P.L.total : 0);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double* lb = lb_all + (size_t)warp * P.lb_doubles;
    const int ngroups = (P.B + P.warps - 1) / P.warps;
    uint32_t phase = 0;
    if (SMEM && threadIdx.x == 0) {
        mbar_init(&cs->bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (;;) {
        __syncthreads();   // previous group fully done with the scenario blocks / cs->work
        if (threadIdx.x == 0) cs->work = atomicAdd(counter, 1);
        __syncthreads();
        const int g = cs->work;
        if (g >= ngroups) break;
        const int b0 = g * P.warps;
        const int blast = min(b0 + P.warps, P.B) - 1;
        const int sc0 = b0 / P.starts, sc1 = blast / P.starts;
        if (SMEM) load_scenarios(P, staged, scn, sc0, sc1, &cs->bar, phase);
        const int b = b0 + warp;
        if (b < P.B) {
            const int sc = b / P.starts;
            const double* S = SMEM ? scn + (size_t)(sc - sc0) * P.L.total : staged + (size_t)sc * P.L.total;
            solve_worker<SPL, 0, 0>(P, S, nullptr, nullptr, lb, b, lane, io);
        }
    }
}

// K1 (queue variant): every warp pulls its own next instance (own CTA's queue first, then the
// others': see solve_worker), so a slow instance never holds other warps at a CTA barrier;
// scenario blocks are read from the staged copy in global memory through L1 (with culling a
// solve touches a few KB of its block, and a CTA works on two or three blocks at a time).
template <int SPL, int FIXED>
__global__ void __launch_bounds__(qthreads(FIXED), MPCB_MIN_CTAS) solve_kernel_queue(const KParams P, const double* __restrict__ staged,
                                                          const SolveIO io, int* __restrict__ counter)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* lb_all = reinterpret_cast<double*>(smem_raw);
    int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    asm volatile("" : "+r"(lane));   // keep the lane id in a register (S2R is slow to re-read)
    double* lb = lb_all + (size_t)warp * P.lb_doubles;
    if (threadIdx.x == 0 && P.prof) {
        unsigned long long tns;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tns));
        P.prof[blockIdx.x & (MPCB_WS_PROF_CTAS - 1)] = tns;
    }
    solve_worker<SPL, 1, FIXED>(P, nullptr, staged, counter, lb, 0, lane, io);
}

// K1 (latency variant, small batches): one CTA per instance, warp 0 solves, the helper warps evaluate
// the line-search trials concurrently (mpcb_solver.cuh "speculative line search").  Same bits as the
// queue kernel.
template <int SPL, int FIXED>
__global__ void __launch_bounds__(SPEC_THREADS, SPL == 1 ? SPEC_CTAS_PER_SM : 1) solve_kernel_spec(const KParams P, const double* __restrict__ staged,
                                                                     const SolveIO io, int* __restrict__ counter)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* lb = reinterpret_cast<double*>(smem_raw);
    SpecShared* SP = reinterpret_cast<SpecShared*>(lb + P.lb_doubles);
    int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0 && P.prof) {
        unsigned long long tns;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tns));
        P.prof[blockIdx.x & (MPCB_WS_PROF_CTAS - 1)] = tns;
    }
    if (warp == 0) {
        asm volatile("" : "+r"(lane));
        solve_worker<SPL, 1, FIXED, false, true>(P, nullptr, staged, counter, lb, 0, lane, io, nullptr, SP);
        if (lane == 0) { SP->cmd = 0; SP->la_cmd = 0; }
        __syncwarp();
        spec_bar<1>();
        spec_bar<3>();
    } else if (warp <= MPCB_SPEC_TRIALS) {
        spec_helper<SPL, FIXED>(P, SP, warp - 1, lane);
    } else {
        spec_lookahead<SPL, FIXED>(P, SP, lane);
    }
}

// K1 (team variant, dimension sets with many ellipses): TEAM_NS solver warps per CTA, each running
// the solve of one instance (same code as the queue kernel, one queue per CTA), share the worker
// warps, which evaluate the per-step part of every horizon evaluation (mpcb_device.cuh "team mode").
template <int SPL, int FIXED>
__global__ void __launch_bounds__(TEAM_THREADS, MPCB_TEAM_CTAS) solve_kernel_team(const KParams P, const double* __restrict__ staged,
                                                                     const SolveIO io, int* __restrict__ counter)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* base = reinterpret_cast<double*>(smem_raw);
    const int sstride = P.lb_doubles + team_solver_doubles(P.L.N);
    TeamPool* pool = reinterpret_cast<TeamPool*>(base + (size_t)TEAM_NS * sstride);
    int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x < TEAM_NS) {
        TeamShared* Ts = reinterpret_cast<TeamShared*>(base + (size_t)threadIdx.x * sstride + P.lb_doubles);
        Ts->req = 0; Ts->done = 0; Ts->exit_ = 0; Ts->pad[0] = (int)threadIdx.x;   // pad[0]: index of the solver
    }
    if (threadIdx.x == 0 && P.prof) {
        unsigned long long tns;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tns));
        P.prof[blockIdx.x & (MPCB_WS_PROF_CTAS - 1)] = tns;
    }
    __syncthreads();
    if (warp < TEAM_NS) {
#ifdef MPCB_TEAM_REGS_SOLVER
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(MPCB_TEAM_REGS_SOLVER));
#endif
        double* lb = base + (size_t)warp * sstride;
        TeamShared* T = reinterpret_cast<TeamShared*>(lb + P.lb_doubles);
        asm volatile("" : "+r"(lane));
        solve_worker<SPL, 1, FIXED, true>(P, nullptr, staged, counter, lb, 0, lane, io, T);
        __syncwarp();
        if (lane == 0) { __threadfence_block(); T->exit_ = 1; }
    } else {
#ifdef MPCB_TEAM_REGS_WORKER
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(MPCB_TEAM_REGS_WORKER));
#endif
        team_worker<FIXED>(P, base, sstride, P.lb_doubles, TEAM_NS, pool, (int)threadIdx.x - 32 * TEAM_NS);
    }
}

// K2 (team variant): one CTA per instance, warp 0 evaluates, same division of labour
template <int SPL>
__global__ void __launch_bounds__(TEAM_THREADS, 1) eval_kernel_team(const KParams P, const double* __restrict__ staged,
                                                                    const double* __restrict__ u,
                                                                    const double* __restrict__ y,
                                                                    const double* __restrict__ c, double* f,
                                                                    double* psi, double* grad, double* F1, double* F2)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* base = reinterpret_cast<double*>(smem_raw);
    TeamShared* T = reinterpret_cast<TeamShared*>(base);
    const int N = P.L.N;
    TeamPool* pool = reinterpret_cast<TeamPool*>(base + team_solver_doubles(N));
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) { T->req = 0; T->done = 0; T->exit_ = 0; T->pad[0] = 0; }
    __syncthreads();
    if (warp >= TEAM_NS) { team_worker<0>(P, base, team_solver_doubles(N), 0, 1, pool, (int)threadIdx.x - 32 * TEAM_NS); return; }
    if (warp != 0) return;
    const int b = blockIdx.x;
    const double* S = staged + (size_t)(b / P.starts) * P.L.total;
    double v[SPL], w[SPL], ya[SPL], yw[SPL];
    bool act[SPL];
#pragma unroll
    for (int j = 0; j < SPL; ++j) {
        const int k = lane + 32 * j;
        act[j] = k < N;
        v[j] = 0.0; w[j] = 0.0; ya[j] = 0.0; yw[j] = 0.0;
        if (act[j]) {
            v[j] = u[(size_t)b * 2 * N + 2 * k];
            w[j] = u[(size_t)b * 2 * N + 2 * k + 1];
            if (y) { ya[j] = y[(size_t)b * 2 * N + k]; yw[j] = y[(size_t)b * 2 * N + N + k]; }
        }
    }
    const double cc = c ? c[b] : P.c_init;
#pragma unroll
    for (int j = 0; j < SPL; ++j) { ya[j] = ya[j] / fmax(cc, 1.0); yw[j] = yw[j] / fmax(cc, 1.0); }
    const int n2 = P.L.Ndyn > 0 ? P.L.Ndyn : 1;
    EvalOut<SPL> o;
    eval_psi<SPL, 0, true>(P, S, v, w, cc, ya, yw, true, o, lane, F2 ? F2 + (size_t)b * n2 : nullptr, true, T);
    __syncwarp();
    if (lane == 0) { __threadfence_block(); T->exit_ = 1; }
    if (lane == 0) {
        if (f) f[b] = o.f;
        if (psi) psi[b] = o.psi;
    }
    double vc = S[P.L.o_hdr + H_UM1V], wc = S[P.L.o_hdr + H_UM1W];
#pragma unroll
    for (int j = 0; j < SPL; ++j) {
        const int k = lane + 32 * j;
        double vp = __shfl_up_sync(FULL, v[j], 1), wp = __shfl_up_sync(FULL, w[j], 1);
        if (lane == 0) { vp = vc; wp = wc; }
        if (SPL > 1) { vc = __shfl_sync(FULL, v[j], 31); wc = __shfl_sync(FULL, w[j], 31); }
        if (act[j]) {
            if (grad) {
                grad[(size_t)b * 2 * N + 2 * k] = o.gv[j];
                grad[(size_t)b * 2 * N + 2 * k + 1] = o.gw[j];
            }
            if (F1) {
                F1[(size_t)b * 2 * N + k] = (v[j] - vp) * P.inv_ts;
                F1[(size_t)b * 2 * N + N + k] = (w[j] - wp) * P.inv_ts;
            }
        }
    }
}

// FP64 FMA throughput probe: the denominator of the roofline bench.py reports (the pool's
// MEASURED_PEAKS.json has no FP64 figure).  8 independent DFMA chains per thread.
__global__ void __launch_bounds__(256) fp64_peak_kernel(int iters, double seed, double* sink)
{
    double a0 = seed + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5,
           a6 = a0 + 6, a7 = a0 + 7;
    const double m = 0.999999, c = 1e-9;
    for (int i = 0; i < iters; ++i) {
        a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
        a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
    }
    if (a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7 == 12345.678) *sink = a0;   // keep the chains alive
}

// ------------------------------------------------------------------ host side
int build_layout(const mpcb_dims* d, Lay& L)
{
    if (!d) return MPCB_E_NULL;
    if (d->N < 1 || d->N > MPCB_MAX_N || d->nedge < 1 || d->nedge > MPCB_MAX_EDGE || d->Nother < 0 ||
        d->Nstc < 0 || d->Ndyn < 0 || d->Ndyn > MPCB_MAX_NDYN)
        return MPCB_E_DIMS;
    L = make_lay(d->N, d->Nother, d->Nstc, d->nedge, d->Ndyn);
    return MPCB_OK;
}

int env_int(const char* name, int dflt)
{
    const char* v = getenv(name);
    return v && *v ? atoi(v) : dflt;
}

constexpr size_t WS_COUNTERS = MPCB_WS_COUNTER_BYTES;   // work-queue counters (one per CTA)
// launch profile of the last solve: start time per CTA, finish time per (CTA, warp), globaltimer ns
constexpr size_t WS_PROF = (size_t)MPCB_WS_PROF_CTAS * (1 + MPCB_WS_PROF_WARPS) * 8;
constexpr size_t WS_HEADER = WS_COUNTERS + WS_PROF;
static_assert(WS_HEADER == MPCB_WS_HEADER_BYTES, "include/mpcb.h documents the workspace header");

struct Plan {
    KParams P;
    bool smem;
    int fixed;     // compiled-in dimension set (0: run-time dims)
    int team;      // worker groups of the team kernels (0: one-warp kernels)
    bool spec;     // latency kernel (speculative line search) for small batches
    int spl;
    size_t smem_bytes;
};

int make_plan(const mpcb_dims* d, const mpcb_robot* r, const mpcb_solver_cfg* c, int n_p, int starts,
              bool need_lbfgs, Plan& pl)
{
    if (!d || !r || !c) return MPCB_E_NULL;
    if (n_p < 1 || starts < 1) return MPCB_E_DIMS;
    if (c->lbfgs_mem < 1 || c->lbfgs_mem > MPCB_MAX_LBFGS || c->max_inner < 1 || c->max_outer < 1)
        return MPCB_E_DIMS;
    KParams& P = pl.P;
    memset(&P, 0, sizeof(P));
    int rc = build_layout(d, P.L);
    if (rc) return rc;
    P.ts = r->ts; P.k6 = r->ts / 6.0; P.inv_ts = 1.0 / r->ts;
    P.ds2 = r->vehicle_width * r->vehicle_width;
    P.dsafe = fabs(r->vehicle_width);
    P.cull = env_int("MPCB_CULL", 1);
    P.vmargin = r->vehicle_margin; P.smargin = r->social_margin;
    P.vmin = r->lin_vel_min; P.vmax = r->lin_vel_max; P.wmax = r->ang_vel_max;
    P.amin = r->lin_acc_min; P.amax = r->lin_acc_max; P.wamax = r->ang_acc_max;
    P.tol = c->tolerance; P.tol0 = c->initial_tolerance; P.delta = c->delta_tolerance;
    P.beta = c->inner_tol_update; P.rho = c->penalty_update; P.theta = c->sufficient_decrease;
    P.c_init = c->initial_penalty; P.sy_eps = c->sy_epsilon; P.cb_eps = c->cbfgs_epsilon;
    P.cb_alpha = c->cbfgs_alpha;
    P.max_inner = c->max_inner; P.max_outer = c->max_outer; P.mem = c->lbfgs_mem;
    P.budget = c->max_inner_total > 0 ? c->max_inner_total : 0;
    P.time_ns = c->max_time_us > 0 ? 1000LL * c->max_time_us : 0;
    P.n_p = n_p; P.starts = starts;
    const long long B = (long long)n_p * starts;
    if (B > 0x7fffffffLL / (2 * d->N)) return MPCB_E_DIMS;
    P.B = (int)B;
    pl.spl = d->N <= 32 ? 1 : 2;
    pl.team = team_groups(d->N, d->Ndyn, c->team_mode == 1);
    P.team_G = pl.team;
    const int M = c->lbfgs_mem + 1;
    // per-warp scratch: L-BFGS rows + rho/alpha, then y, y+ and the parked solver state
    P.lb_doubles = need_lbfgs ? (((2 * M * 2 * d->N + 2 * M + 1) & ~1) + scratch_doubles(d->N)) : 0;

    // choose warps per CTA so that the scenario blocks + per-warp L-BFGS fit in
    // shared memory; fall back to reading the staged blocks from global (L1/L2)
    const size_t cap = 227 * 1024;
    const size_t blk = (size_t)P.L.total * 8, lbw = (size_t)P.lb_doubles * 8;
    pl.smem = false;
    pl.fixed = 0;
    int best_wps = 0;
    for (int W = 8; W >= 1; W >>= 1) {
        int nsc;
        if (starts % W == 0) nsc = 1;
        else if (W % starts == 0) nsc = W / starts;
        else nsc = (W + starts - 2) / starts + 1;
        const size_t need = 16 + nsc * blk + W * lbw;
        if (need > cap) continue;
        int ctas = (int)(cap / (need + 1024));   // + per-CTA reserved shared memory
        if (ctas < 1) ctas = 1;
        if (ctas * W > 32) ctas = 32 / W;
        const int wps = ctas * W;                // resident warps per SM
        if (wps > best_wps) {
            best_wps = wps; pl.smem = true; P.warps = W; P.nsc = nsc; pl.smem_bytes = need;
        }
    }
    const int mode = env_int("MPCB_MODE", 0);   // 0 auto (queue), 1 force shared-memory groups
    if (best_wps < 8 || (mode != 1 && need_lbfgs)) {
        // queue variant (solve) / too few resident warps: read the staged blocks through L1/L2
        pl.smem = false;
        pl.fixed = 0;
        if (need_lbfgs && env_int("MPCB_FIXED", 1) && c->lbfgs_mem == MPCB_FIX_MEM)
            for (int fx = 1; fx <= MPCB_NUM_FIXED; ++fx) {
                const Lay F = fixed_lay(fx);
                if (d->N == F.N && d->Nother == F.Nother && d->Nstc == F.Nstc && d->nedge == F.nedge &&
                    d->Ndyn == F.Ndyn)
                    pl.fixed = fx;
            }
        const int maxw = qthreads(pl.fixed) / 32;
        P.warps = env_int("MPCB_WARPS", maxw); P.nsc = 0;
        if (P.warps < 1 || P.warps > maxw) P.warps = maxw < 8 ? maxw : 8;
        while (P.warps > 1 && 16 + P.warps * lbw > cap) --P.warps;   // large N: fewer warps per CTA
        // a batch of fewer instances than resident warps is spread over all SMs (a warp alone on a
        // scheduler runs an iteration 1.5x faster than beside three others) instead of filling a few
        if (need_lbfgs && env_int("MPCB_WARPS", 0) == 0) {
            const long long per_sm = ((long long)B + 147) / 148;
            if (per_sm < P.warps) P.warps = per_sm < 1 ? 1 : (int)per_sm;
        }
        pl.smem_bytes = 16 + P.warps * lbw;
    }
    // small batches: the latency kernel, same bits as the queue kernel.  Its persistent CTAs (two per
    // SM) pull instances from the same queues; measured against the one-warp kernel at default dims
    // (ms per batch, 8 starts per scenario): B = 592: 115 vs 346, 1184: 168 vs 329, 1776: 245 vs 266,
    // 2072: 246 vs 246, 2368: 273 vs 257 - a batch this small is bound by its slowest instance, which a CTA
    // solves twice as fast
    pl.spec = need_lbfgs && !pl.team && !pl.smem &&
              B <= env_int("MPCB_SPEC_MAXB", (pl.spl == 1 ? 12 : 2) * 148) && env_int("MPCB_SPEC", 1) != 0;
    if (pl.spec) pl.smem_bytes = (size_t)(P.lb_doubles + spec_doubles(d->N)) * 8;
    if (pl.team) {
        // per solver warp its scratch and its team block, then the worker pool's scratch; the
        // evaluation kernel has one solver and no solver scratch
        pl.smem = false;
        P.warps = MPCB_TEAM_WARPS; P.nsc = 0;
        pl.smem_bytes = (size_t)team_smem_doubles(d->N, pl.team, P.lb_doubles, need_lbfgs ? TEAM_NS : 1) * 8;
    }
    return MPCB_OK;
}

template <typename K>
int set_smem(K kernel, size_t bytes)
{
    CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    return MPCB_OK;
}

// workspace: header | order[n_p] (int) | keys[n_p] (float) | staged scenario blocks
size_t ws_order_bytes(int n_p) { return ((size_t)n_p * 8 + 255) & ~(size_t)255; }

int stage(Plan& pl, const double* p, void* workspace, size_t ws_bytes, cudaStream_t st,
          double** staged_out, int** counter_out, bool ordered)
{
    const size_t need = WS_HEADER + ws_order_bytes(pl.P.n_p) + (size_t)pl.P.n_p * pl.P.L.total * 8;
    if (!workspace) return MPCB_E_NULL;
    if (ws_bytes < need) return MPCB_E_WORKSPACE;
    if ((reinterpret_cast<uintptr_t>(workspace) & 15) || (reinterpret_cast<uintptr_t>(p) & 7)) return MPCB_E_ALIGN;
    int* counter = reinterpret_cast<int*>(workspace);
    int* order = reinterpret_cast<int*>(reinterpret_cast<char*>(workspace) + WS_HEADER);
    float* keys = reinterpret_cast<float*>(order + pl.P.n_p);
    double* staged = reinterpret_cast<double*>(reinterpret_cast<char*>(workspace) + WS_HEADER + ws_order_bytes(pl.P.n_p));
    CUDA_TRY(cudaMemsetAsync(counter, 0, WS_HEADER, st));   // counters and launch profile
    const int grid = pl.P.n_p < 148 * 16 ? pl.P.n_p : 148 * 16;
    ordered = ordered && env_int("MPCB_ORDER", 1) != 0;
    stage_kernel<<<grid, 128, 0, st>>>(pl.P, p, staged, ordered ? keys : nullptr);
    CUDA_TRY(cudaGetLastError());
    pl.P.order = nullptr;
    if (ordered) {
        order_kernel<<<1, 1024, 0, st>>>(pl.P.n_p, keys, order);
        CUDA_TRY(cudaGetLastError());
        pl.P.order = order;
    }
    *staged_out = staged;
    *counter_out = counter;
    return MPCB_OK;
}

}  // namespace

extern "C" {

int32_t mpcb_abi_version(void) { return MPCB_ABI_VERSION; }
const char* mpcb_last_error(void) { return g_err; }

int32_t mpcb_param_len(const mpcb_dims* d)
{
    Lay L;
    return build_layout(d, L) ? -1 : L.np;
}
int32_t mpcb_num_decision(const mpcb_dims* d) { return d ? 2 * d->N : -1; }
int32_t mpcb_n1(const mpcb_dims* d) { return d ? 2 * d->N : -1; }
int32_t mpcb_n2(const mpcb_dims* d) { return d ? (d->Ndyn > 0 ? d->Ndyn : 1) : -1; }
int32_t mpcb_team_groups(const mpcb_dims* d) { return d ? team_groups(d->N, d->Ndyn) : -1; }
int32_t mpcb_team_groups_cfg(const mpcb_dims* d, const mpcb_solver_cfg* c)
{
    return d && c ? team_groups(d->N, d->Ndyn, c->team_mode == 1) : -1;
}

void mpcb_default_robot(mpcb_robot* r)
{
    if (!r) return;
    r->ts = 0.2; r->vehicle_width = 0.5; r->vehicle_margin = 0.2; r->social_margin = 0.2;
    r->lin_vel_min = -0.5; r->lin_vel_max = 1.5; r->ang_vel_max = 0.5;
    r->lin_acc_min = -1.0; r->lin_acc_max = 1.0; r->ang_acc_max = 3.0;
}
void mpcb_default_solver_cfg(mpcb_solver_cfg* c)
{
    if (!c) return;
    c->tolerance = 1e-4; c->initial_tolerance = 1e-4; c->delta_tolerance = 1e-4;
    c->inner_tol_update = 0.1; c->penalty_update = 5.0; c->sufficient_decrease = 0.1;
    c->initial_penalty = 10.0; c->sy_epsilon = 1e-10; c->cbfgs_epsilon = 1e-8; c->cbfgs_alpha = 1.0;
    c->max_inner = 500; c->max_outer = 10; c->lbfgs_mem = 10; c->max_inner_total = 0;
    c->team_mode = 0; c->max_time_us = 0;
}

int32_t mpcb_workspace_bytes(const mpcb_dims* d, int32_t n_p, int32_t starts, size_t* bytes)
{
    Lay L;
    int rc = build_layout(d, L);
    if (rc) return rc;
    if (!bytes) return MPCB_E_NULL;
    if (n_p < 1 || starts < 1) return MPCB_E_DIMS;
    *bytes = WS_HEADER + ws_order_bytes(n_p) + (size_t)n_p * L.total * 8;
    return MPCB_OK;
}

int32_t mpcb_eval_f64(const mpcb_dims* d, const mpcb_robot* r, const mpcb_solver_cfg* c, int32_t n_p,
                      int32_t starts, const double* p, const double* u, const double* y,
                      const double* cpen, double* f, double* psi, double* grad, double* F1, double* F2,
                      void* workspace, size_t ws_bytes, void* stream)
{
    if (!p || !u) return MPCB_E_NULL;
    Plan pl;
    int rc = make_plan(d, r, c, n_p, starts, false, pl);
    if (rc) return rc;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    double* staged; int* counter;
    rc = stage(pl, p, workspace, ws_bytes, st, &staged, &counter, false);
    if (rc) return rc;
    if (pl.team) {
        if (pl.spl == 1) {
            rc = set_smem(eval_kernel_team<1>, pl.smem_bytes);
            if (rc) return rc;
            eval_kernel_team<1><<<pl.P.B, TEAM_THREADS, pl.smem_bytes, st>>>(pl.P, staged, u, y, cpen, f, psi, grad, F1, F2);
        } else {
            rc = set_smem(eval_kernel_team<2>, pl.smem_bytes);
            if (rc) return rc;
            eval_kernel_team<2><<<pl.P.B, TEAM_THREADS, pl.smem_bytes, st>>>(pl.P, staged, u, y, cpen, f, psi, grad, F1, F2);
        }
        CUDA_TRY(cudaGetLastError());
        return MPCB_OK;
    }
    const int grid = (pl.P.B + pl.P.warps - 1) / pl.P.warps;
    const int threads = pl.P.warps * 32;
#define LAUNCH_EVAL(SPL, SM)                                                                       \
    do {                                                                                           \
        rc = set_smem(eval_kernel<SPL, SM>, pl.smem_bytes);                                        \
        if (rc) return rc;                                                                         \
        eval_kernel<SPL, SM><<<grid, threads, pl.smem_bytes, st>>>(pl.P, staged, u, y, cpen, f,    \
                                                                    psi, grad, F1, F2);            \
    } while (0)
    if (pl.spl == 1) { if (pl.smem) LAUNCH_EVAL(1, true); else LAUNCH_EVAL(1, false); }
    else             { if (pl.smem) LAUNCH_EVAL(2, true); else LAUNCH_EVAL(2, false); }
#undef LAUNCH_EVAL
    CUDA_TRY(cudaGetLastError());
    return MPCB_OK;
}

int32_t mpcb_solve_f64(const mpcb_dims* d, const mpcb_robot* r, const mpcb_solver_cfg* c, int32_t n_p,
                       int32_t starts, const double* p, const double* u0, const double* y0,
                       const double* c0, double* u_out, double* cost, int32_t* exit_status,
                       int32_t* n_outer, int32_t* n_inner, double* fpr, double* f1_infeas,
                       double* f2_norm, double* penalty, double* y_out, int32_t* evals,
                       void* workspace, size_t ws_bytes, void* stream)
{
    if (!p || !u_out || !exit_status) return MPCB_E_NULL;
    if ((reinterpret_cast<uintptr_t>(u_out) & 15) || (u0 && (reinterpret_cast<uintptr_t>(u0) & 15)))
        return MPCB_E_ALIGN;
    Plan pl;
    int rc = make_plan(d, r, c, n_p, starts, true, pl);
    if (rc) return rc;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    double* staged; int* counter;
    // hardest-first order: pays when the launch is only a few rounds of instances per resident warp
    // (+6 % at configs[2], 2.3 instances per warp); with dozens of rounds (the headline workload) the
    // key is too coarse a predictor to shorten the tail (measured: no gain), so the natural order stays
    const bool few_rounds = (long long)n_p * starts <= 8LL * 148 * 16;
    rc = stage(pl, p, workspace, ws_bytes, st, &staged, &counter, n_p > 1 && (few_rounds || env_int("MPCB_ORDER", 1) == 2));
    if (rc) return rc;
    SolveIO io{u0, y0, c0, u_out, cost, exit_status, n_outer, n_inner, fpr, f1_infeas, f2_norm, penalty, y_out, evals};
    pl.P.prof = reinterpret_cast<unsigned long long*>(reinterpret_cast<char*>(workspace) + WS_COUNTERS);
    const int threads = pl.P.warps * 32;
    const int ngroups = (pl.P.B + pl.P.warps - 1) / pl.P.warps;
    int dev = 0, sms = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
#define LAUNCH_SOLVE(SPL, SM)                                                                      \
    do {                                                                                           \
        rc = set_smem(solve_kernel<SPL, SM>, pl.smem_bytes);                                       \
        if (rc) return rc;                                                                         \
        int per_sm = 0;                                                                            \
        CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, solve_kernel<SPL, SM>,     \
                                                               threads, pl.smem_bytes));           \
        if (per_sm < 1) per_sm = 1;                                                                \
        int grid = sms * per_sm;                                                                   \
        if (grid > ngroups) grid = ngroups;                                                        \
        solve_kernel<SPL, SM><<<grid, threads, pl.smem_bytes, st>>>(pl.P, staged, io, counter);    \
    } while (0)
#define LAUNCH_QUEUE(SPL, MD)                                                                      \
    do {                                                                                           \
        rc = set_smem(solve_kernel_queue<SPL, MD>, pl.smem_bytes);                                 \
        if (rc) return rc;                                                                         \
        int per_sm = 0;                                                                            \
        CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, solve_kernel_queue<SPL, MD>, \
                                                               threads, pl.smem_bytes));           \
        if (per_sm < 1) per_sm = 1;                                                                \
        if (env_int("MPCB_CTAS_PER_SM", 0) > 0) per_sm = env_int("MPCB_CTAS_PER_SM", 0);           \
        int grid = sms * per_sm;                                                                   \
        if (grid > ngroups) grid = ngroups;                                                        \
        if (grid > (int)(WS_COUNTERS / sizeof(int))) grid = (int)(WS_COUNTERS / sizeof(int));      \
        solve_kernel_queue<SPL, MD><<<grid, threads, pl.smem_bytes, st>>>(pl.P, staged, io, counter); \
    } while (0)
#define LAUNCH_TEAM(SPL, MD)                                                                       \
    do {                                                                                           \
        rc = set_smem(solve_kernel_team<SPL, MD>, pl.smem_bytes);                                  \
        if (rc) return rc;                                                                         \
        int per_sm = 0;                                                                            \
        CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, solve_kernel_team<SPL, MD>, \
                                                               TEAM_THREADS, pl.smem_bytes));      \
        if (per_sm < 1) per_sm = 1;                                                                \
        int grid = sms * per_sm;                                                                   \
        if (grid > (pl.P.B + TEAM_NS - 1) / TEAM_NS) grid = (pl.P.B + TEAM_NS - 1) / TEAM_NS;       \
        if (grid > (int)(WS_COUNTERS / sizeof(int))) grid = (int)(WS_COUNTERS / sizeof(int));      \
        solve_kernel_team<SPL, MD><<<grid, TEAM_THREADS, pl.smem_bytes, st>>>(pl.P, staged, io, counter); \
    } while (0)
#define LAUNCH_SPEC(SPL, MD)                                                                       \
    do {                                                                                           \
        rc = set_smem(solve_kernel_spec<SPL, MD>, pl.smem_bytes);                                  \
        if (rc) return rc;                                                                         \
        int grid = pl.P.B < 2 * sms ? pl.P.B : 2 * sms;                                            \
        solve_kernel_spec<SPL, MD><<<grid, SPEC_THREADS, pl.smem_bytes, st>>>(pl.P, staged, io, counter); \
    } while (0)
    if (pl.team) {
        if (pl.fixed == 3) LAUNCH_TEAM(2, 3);
        else if (pl.spl == 1) LAUNCH_TEAM(1, 0);
        else LAUNCH_TEAM(2, 0);
    }
    else if (pl.spec) {
        if (pl.fixed == 1) LAUNCH_SPEC(1, 1);
        else if (pl.spl == 1) LAUNCH_SPEC(1, 0);
        else LAUNCH_SPEC(2, 0);
    }
    else if (pl.smem) { if (pl.spl == 1) LAUNCH_SOLVE(1, true); else LAUNCH_SOLVE(2, true); }
    else {
        if (pl.fixed == 1) LAUNCH_QUEUE(1, 1);
        else if (pl.fixed == 2) LAUNCH_QUEUE(1, 2);
        else if (pl.spl == 1) LAUNCH_QUEUE(1, 0);
        else LAUNCH_QUEUE(2, 0);
    }
#undef LAUNCH_SOLVE
#undef LAUNCH_QUEUE
#undef LAUNCH_TEAM
#undef LAUNCH_SPEC
    CUDA_TRY(cudaGetLastError());
    return MPCB_OK;
}

// Per-thread context of the single-solve host path: one device arena, one pinned staging buffer,
// a private non-blocking stream and two events, all released when the thread ends.
struct OneHostCtx {
    char* dev = nullptr; char* pin = nullptr; size_t bytes = 0, pin_bytes = 0;
    cudaStream_t st = nullptr; cudaEvent_t e0 = nullptr, e1 = nullptr;
    ~OneHostCtx()
    {   // best effort: at process exit the CUDA context may already be gone
        if (dev) cudaFree(dev);
        if (pin) cudaFreeHost(pin);
        if (e0) cudaEventDestroy(e0);
        if (e1) cudaEventDestroy(e1);
        if (st) cudaStreamDestroy(st);
    }
};

int32_t mpcb_solve_one_host(const mpcb_dims* d, const mpcb_robot* r, const mpcb_solver_cfg* c,
                            const double* p_host, const double* u0_host, const double* y0_host,
                            const double* c0_host, double* u_out_host, double* y_out_host,
                            int32_t* exit_status_host, double* out_scalars)
{
    if (!p_host || !u_out_host || !exit_status_host) return MPCB_E_NULL;
    Lay L;
    int rc = build_layout(d, L);
    if (rc) return rc;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) {
        snprintf(g_err, sizeof(g_err), "no CUDA device");
        return MPCB_E_NO_DEVICE;
    }
    const int n = 2 * d->N;
    size_t ws = 0;
    mpcb_workspace_bytes(d, 1, 1, &ws);
    // one arena, mirrored on the device and in pinned host memory; every block starts on a
    // 256-byte boundary (u0 / u_out feed double2 loads and stores: any np, odd ones included):
    //   inputs  p | u0 | y0 | c0     outputs  u | y | scalars(5 doubles) | ints(7)     workspace
    auto up = [](size_t x) { return (x + 255) & ~(size_t)255; };
    const size_t o_p = 0, o_u0 = up(o_p + (size_t)L.np * 8), o_y0 = up(o_u0 + (size_t)n * 8),
                 o_c0 = up(o_y0 + (size_t)n * 8), o_in_end = up(o_c0 + 8);
    const size_t o_u = o_in_end, o_y = up(o_u + (size_t)n * 8), o_sc = up(o_y + (size_t)n * 8),
                 o_i = up(o_sc + 5 * 8), o_out_end = up(o_i + 7 * 4), o_ws = o_out_end;
    // per-thread cached context: the single-solve path is called once per control period, so
    // allocation must not be on it
    static thread_local OneHostCtx t;
    if (t.bytes < o_ws + ws) {
        if (t.dev) cudaFree(t.dev);
        t.dev = nullptr; t.bytes = 0;
        CUDA_TRY(cudaMalloc(&t.dev, o_ws + ws));
        t.bytes = o_ws + ws;
    }
    if (t.pin_bytes < o_out_end) {
        if (t.pin) cudaFreeHost(t.pin);
        t.pin = nullptr; t.pin_bytes = 0;
        CUDA_TRY(cudaMallocHost(&t.pin, o_out_end));
        t.pin_bytes = o_out_end;
    }
    if (!t.st) {
        CUDA_TRY(cudaStreamCreateWithFlags(&t.st, cudaStreamNonBlocking));
        CUDA_TRY(cudaEventCreate(&t.e0));
        CUDA_TRY(cudaEventCreate(&t.e1));
    }
    char* arena = t.dev;
    cudaStream_t st = t.st;
    // stage the inputs in pinned memory: one asynchronous copy instead of pageable ones
    memcpy(t.pin + o_p, p_host, (size_t)L.np * 8);
    if (u0_host) memcpy(t.pin + o_u0, u0_host, (size_t)n * 8);
    if (y0_host) memcpy(t.pin + o_y0, y0_host, (size_t)n * 8);
    if (c0_host) memcpy(t.pin + o_c0, c0_host, 8);
    CUDA_TRY(cudaMemcpyAsync(arena, t.pin, o_in_end, cudaMemcpyHostToDevice, st));
    double* sc = reinterpret_cast<double*>(arena + o_sc);
    int32_t* iv = reinterpret_cast<int32_t*>(arena + o_i);
    CUDA_TRY(cudaEventRecord(t.e0, st));
    rc = mpcb_solve_f64(d, r, c, 1, 1, reinterpret_cast<double*>(arena + o_p),
                        u0_host ? reinterpret_cast<double*>(arena + o_u0) : nullptr,
                        y0_host ? reinterpret_cast<double*>(arena + o_y0) : nullptr,
                        c0_host ? reinterpret_cast<double*>(arena + o_c0) : nullptr,
                        reinterpret_cast<double*>(arena + o_u), sc + 0, iv + 0, iv + 1, iv + 2, sc + 1,
                        sc + 2, sc + 3, sc + 4, reinterpret_cast<double*>(arena + o_y), iv + 3,
                        arena + o_ws, ws, st);
    if (rc) { cudaStreamSynchronize(st); return rc; }
    CUDA_TRY(cudaEventRecord(t.e1, st));
    CUDA_TRY(cudaMemcpyAsync(t.pin + o_u, arena + o_u, o_out_end - o_u, cudaMemcpyDeviceToHost, st));
    const cudaError_t e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) {
        snprintf(g_err, sizeof(g_err), "solve: %s", cudaGetErrorString(e));
        return MPCB_E_CUDA;
    }
    float ms = 0.f;
    CUDA_TRY(cudaEventElapsedTime(&ms, t.e0, t.e1));
    memcpy(u_out_host, t.pin + o_u, (size_t)n * 8);
    if (y_out_host) memcpy(y_out_host, t.pin + o_y, (size_t)n * 8);
    const double* hsc = reinterpret_cast<const double*>(t.pin + o_sc);
    const int32_t* hiv = reinterpret_cast<const int32_t*>(t.pin + o_i);
    *exit_status_host = hiv[0];
    if (out_scalars) {
        out_scalars[0] = hsc[0]; out_scalars[1] = hsc[1]; out_scalars[2] = hsc[2];
        out_scalars[3] = hsc[3]; out_scalars[4] = hsc[4];
        out_scalars[5] = hiv[1]; out_scalars[6] = hiv[2]; out_scalars[7] = ms;
    }
    return MPCB_OK;
}

// ---- closed-loop support --------------------------------------------
void mpcb_sincos_host(double x, double* sn, double* cs) { mpcb::sincos_cw(x, sn, cs); }

int32_t mpcb_pack_f64(const mpcb_dims* d, const mpcb_sim* sim, double* p_out, void* stream)
{
    if (!d || !sim || !p_out) return MPCB_E_NULL;
    Lay L;
    int rc = build_layout(d, L);
    if (rc) return rc;
    if (sim->n < 1 || sim->T < 1 || sim->Kp < 0 || sim->Kp > 64 || sim->Pd < 0 || sim->M < 1) return MPCB_E_DIMS;
    if (!sim->state || !sim->last_u || !sim->ref_traj || !sim->ref_len || !sim->idx_ref || !sim->goal ||
        !sim->n_poly || !sim->done || (sim->Kp && !sim->polys) || (sim->Pd && (!sim->ped_pos || !sim->ped_vel)))
        return MPCB_E_NULL;
    SimArgs A;
    static_assert(sizeof(SimArgs) == sizeof(mpcb_sim), "mpcb_sim must mirror SimArgs");
    memcpy(&A, sim, sizeof(A));
    pack_kernel<<<sim->n, 64, 0, reinterpret_cast<cudaStream_t>(stream)>>>(L, A, p_out);
    CUDA_TRY(cudaGetLastError());
    return MPCB_OK;
}

int32_t mpcb_plant_step_f64(const mpcb_dims* d, const mpcb_sim* sim, const double* u, void* stream)
{
    if (!d || !sim || !u) return MPCB_E_NULL;
    if (sim->n < 1 || d->N < 1) return MPCB_E_DIMS;
    SimArgs A;
    memcpy(&A, sim, sizeof(A));
    plant_kernel<<<(sim->n + 127) / 128, 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(A, 2 * d->N, u);
    CUDA_TRY(cudaGetLastError());
    return MPCB_OK;
}

int32_t mpcb_cluster_f64(const mpcb_dims* d, int32_t n, int32_t K, int32_t H, double eps, int32_t min_samples,
                         double enlarge, double human_size, const double* hyp, const int32_t* n_hyp,
                         const double* cur_pos, double* o_d, int32_t* scratch, void* stream)
{
    if (!d || !hyp || !cur_pos || !o_d || !scratch) return MPCB_E_NULL;
    if (n < 1 || K < 1 || K > CL_MAXK || H < 0 || d->N < 1 || d->Ndyn < 1 || min_samples < 1) return MPCB_E_DIMS;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const int total = n * (d->N + 1);
    cluster_kernel<<<(total + 63) / 64, 64, 0, st>>>(n, d->N, K, H, d->Ndyn, eps, min_samples, enlarge,
                                                     human_size, hyp, n_hyp, cur_pos, o_d, scratch);
    cluster_fill_kernel<<<n, 128, 0, st>>>(n, d->N, d->Ndyn, scratch, o_d);
    CUDA_TRY(cudaGetLastError());
    return MPCB_OK;
}

/* ---- f32 twins: single-precision buffers at the boundary ------------------------
 * The arithmetic stays f64.  PANOC's Lipschitz probe perturbs the iterate by 1e-12 (below f32
 * resolution) and the 1e-6 radius regulariser makes the ellipse terms reach 1e12 x distance^2
 * (SURVEY Appendix B-2, C-11): neither survives single precision, so the f32 mode is a boundary
 * mode - parameters, guesses and results travel as float32 (half the host<->device bytes), are
 * widened exactly on the device, solved by the f64 kernels, and the results rounded once.  The
 * outputs therefore equal the f64 entry point's on the widened inputs, rounded to f32. */
namespace {
__global__ void widen_kernel(const float* __restrict__ a, double* __restrict__ b, size_t n)
{
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        b[i] = (double)a[i];
}
__global__ void narrow_kernel(const double* __restrict__ a, float* __restrict__ b, size_t n)
{
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        b[i] = (float)a[i];
}
inline size_t up256(size_t x) { return (x + 255) & ~(size_t)255; }
struct F32Layout {
    size_t ws64, p, u0, y0, c0, u, y, sc, total;   // byte offsets inside the workspace; sc: 5 x [B] scalars
};
int f32_layout(const mpcb_dims* d, int32_t n_p, int32_t starts, F32Layout& F)
{
    Lay L;
    int rc = build_layout(d, L);
    if (rc) return rc;
    if (n_p < 1 || starts < 1) return MPCB_E_DIMS;
    size_t ws = 0;
    rc = mpcb_workspace_bytes(d, n_p, starts, &ws);
    if (rc) return rc;
    const size_t B = (size_t)n_p * starts, n = 2 * (size_t)d->N;
    F.ws64 = up256(ws);
    F.p = F.ws64;
    F.u0 = F.p + up256((size_t)n_p * L.np * 8);
    F.y0 = F.u0 + up256(B * n * 8);
    F.c0 = F.y0 + up256(B * n * 8);
    F.u = F.c0 + up256(B * 8);
    F.y = F.u + up256(B * n * 8);
    F.sc = F.y + up256(B * n * 8);
    F.total = F.sc + 5 * up256(B * 8);
    return MPCB_OK;
}
void widen(const float* a, void* base, size_t off, size_t n, cudaStream_t st)
{
    if (a && n) widen_kernel<<<(unsigned)((n + 255) / 256 < 4096 ? (n + 255) / 256 : 4096), 256, 0, st>>>(
        a, reinterpret_cast<double*>(reinterpret_cast<char*>(base) + off), n);
}
void narrow(void* base, size_t off, float* b, size_t n, cudaStream_t st)
{
    if (b && n) narrow_kernel<<<(unsigned)((n + 255) / 256 < 4096 ? (n + 255) / 256 : 4096), 256, 0, st>>>(
        reinterpret_cast<const double*>(reinterpret_cast<char*>(base) + off), b, n);
}
}  // namespace

int32_t mpcb_workspace_bytes_f32(const mpcb_dims* d, int32_t n_p, int32_t starts, size_t* bytes)
{
    if (!bytes) return MPCB_E_NULL;
    F32Layout F;
    int rc = f32_layout(d, n_p, starts, F);
    if (rc) return rc;
    *bytes = F.total;
    return MPCB_OK;
}

int32_t mpcb_solve_f32(const mpcb_dims* d, const mpcb_robot* r, const mpcb_solver_cfg* c, int32_t n_p,
                       int32_t starts, const float* p, const float* u0, const float* y0, const float* c0,
                       float* u_out, float* cost, int32_t* exit_status, int32_t* n_outer, int32_t* n_inner,
                       float* fpr, float* f1_infeas, float* f2_norm, float* penalty, float* y_out,
                       int32_t* evals, void* workspace, size_t ws_bytes, void* stream)
{
    if (!p || !u_out || !exit_status || !workspace) return MPCB_E_NULL;
    F32Layout F;
    int rc = f32_layout(d, n_p, starts, F);
    if (rc) return rc;
    if (ws_bytes < F.total) return MPCB_E_WORKSPACE;
    if (reinterpret_cast<uintptr_t>(workspace) & 255) return MPCB_E_ALIGN;
    Lay L;
    build_layout(d, L);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const size_t B = (size_t)n_p * starts, n = 2 * (size_t)d->N;
    char* w = reinterpret_cast<char*>(workspace);
    auto dp = [&](size_t off) { return reinterpret_cast<double*>(w + off); };
    widen(p, w, F.p, (size_t)n_p * L.np, st);
    widen(u0, w, F.u0, B * n, st);
    widen(y0, w, F.y0, B * n, st);
    widen(c0, w, F.c0, B, st);
    const size_t s1 = up256(B * 8);
    rc = mpcb_solve_f64(d, r, c, n_p, starts, dp(F.p), u0 ? dp(F.u0) : nullptr, y0 ? dp(F.y0) : nullptr,
                        c0 ? dp(F.c0) : nullptr, dp(F.u), dp(F.sc), exit_status, n_outer, n_inner, dp(F.sc + s1),
                        dp(F.sc + 2 * s1), dp(F.sc + 3 * s1), dp(F.sc + 4 * s1), dp(F.y), evals, workspace, F.ws64, stream);
    if (rc) return rc;
    narrow(w, F.u, u_out, B * n, st);
    narrow(w, F.y, y_out, B * n, st);
    narrow(w, F.sc, cost, B, st);
    narrow(w, F.sc + s1, fpr, B, st);
    narrow(w, F.sc + 2 * s1, f1_infeas, B, st);
    narrow(w, F.sc + 3 * s1, f2_norm, B, st);
    narrow(w, F.sc + 4 * s1, penalty, B, st);
    CUDA_TRY(cudaGetLastError());
    return MPCB_OK;
}

int32_t mpcb_eval_f32(const mpcb_dims* d, const mpcb_robot* r, const mpcb_solver_cfg* c, int32_t n_p,
                      int32_t starts, const float* p, const float* u, const float* y, const float* cpen,
                      float* f, float* psi, float* grad, float* F1, float* F2, void* workspace, size_t ws_bytes,
                      void* stream)
{
    if (!p || !u || !workspace) return MPCB_E_NULL;
    F32Layout F;
    int rc = f32_layout(d, n_p, starts, F);
    if (rc) return rc;
    const size_t B = (size_t)n_p * starts, n = 2 * (size_t)d->N, n2 = d->Ndyn > 0 ? d->Ndyn : 1;
    // F2 [B, n2] does not fit the solve layout's scalar slots in general: it follows the layout
    const size_t oF2 = F.total, need = F.total + up256(B * n2 * 8);
    if (ws_bytes < need) return MPCB_E_WORKSPACE;
    if (reinterpret_cast<uintptr_t>(workspace) & 255) return MPCB_E_ALIGN;
    Lay L;
    build_layout(d, L);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    char* w = reinterpret_cast<char*>(workspace);
    auto dp = [&](size_t off) { return reinterpret_cast<double*>(w + off); };
    widen(p, w, F.p, (size_t)n_p * L.np, st);
    widen(u, w, F.u0, B * n, st);
    widen(y, w, F.y0, B * n, st);
    widen(cpen, w, F.c0, B, st);
    const size_t s1 = up256(B * 8);
    // slots: f -> sc[0], psi -> sc[1], grad -> u, F1 -> y, F2 -> after the layout
    rc = mpcb_eval_f64(d, r, c, n_p, starts, dp(F.p), dp(F.u0), y ? dp(F.y0) : nullptr, cpen ? dp(F.c0) : nullptr,
                       dp(F.sc), dp(F.sc + s1), dp(F.u), dp(F.y), dp(oF2), workspace, F.ws64, stream);
    if (rc) return rc;
    narrow(w, F.sc, f, B, st);
    narrow(w, F.sc + s1, psi, B, st);
    narrow(w, F.u, grad, B * n, st);
    narrow(w, F.y, F1, B * n, st);
    narrow(w, oF2, F2, B * n2, st);
    CUDA_TRY(cudaGetLastError());
    return MPCB_OK;
}

/* Measured FP64 FMA throughput of the current device in TFLOP/s (2 flop per FMA). */
int32_t mpcb_fp64_peak_tflops(double* tflops)
{
    if (!tflops) return MPCB_E_NULL;
    int dev = 0, sms = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    double* sink = nullptr;
    CUDA_TRY(cudaMalloc(&sink, 8));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 1 << 15, grid = sms * 8;
    fp64_peak_kernel<<<grid, 256>>>(1024, 1.0, sink);       // warm-up
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) {
        cudaEventRecord(e0);
        fp64_peak_kernel<<<grid, 256>>>(iters, 1.0, sink);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        best = ms < best ? ms : best;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(sink);
    CUDA_TRY(cudaGetLastError());
    *tflops = 2.0 * 8.0 * iters * 256.0 * grid / (best * 1e-3) / 1e12;
    return MPCB_OK;
}

}  // extern "C"
