// mpcb_sim.cuh — device-side packer and plant step (f-2).
//
// K4 pack_kernel: one CTA per episode builds its parameter row p, in the reference layout
//   (mpc_builder.py:47-60), the way the reference's host code does every timestep:
//   reference window    TrajectoryTracker.get_ref_states        (trajectory_tracker.py:243-270),
//                       searched in [idx - N_hor, idx + 5 N_hor) as the reference's call does (:187)
//   speed reference     run_step incl. its max() quirk          (:304-310)
//   o_d                 MainBase.run_one_step ellipse list      (main_base.py:293-302) from
//                       constant-velocity pedestrian modes, MpcInterface.get_dyn_constraints
//   o_s                 the Nstcobs polygons closest by edge distance -> half-spaces
//                       (mpc_interface.py:73-100, utils_geo.py:6-62)
//   concatenation       trajectory_tracker.py:315-317
// K5 plant_kernel: RK4 unicycle step with the first action (motion_model.py:141-163),
//   pedestrians advance, termination test (trajectory_tracker.py:191-199).
// Arithmetic mirrors dyobav_mpcnwta_warehouse_b200/closed_loop.py operation for operation
// (no fused multiply-adds except inside sincos_cw), so host and device loops agree bitwise.
#pragma once
#include "mpcb_device.cuh"

namespace mpcb {

struct SimArgs {
    int n;                 // episodes
    int T;                 // padded length of every reference trajectory
    int Kp;                // polygon slots per episode (4 vertices each)
    int Pd, M;             // pedestrians per episode, modes per pedestrian
    double base_speed, lin_vel_max, ped_size, stc_w, dyn_w, ts;
    double tuning[10];
    // per-episode state (device pointers)
    double* state;         // [n,3]
    double* last_u;        // [n,2]
    const double* ref_traj;  // [n,T,3]
    const int* ref_len;    // [n]
    int* idx_ref;          // [n]
    const double* goal;    // [n,2]
    const double* polys;   // [n,Kp,4,2]
    const int* n_poly;     // [n]
    double* ped_pos;       // [n,Pd,2]
    const double* ped_vel; // [n,Pd,M,2]
    int* done;             // [n]
    const double* od_in;   // [n,Ndyn,N+1,6] or NULL: the o_d block as a predictor stage produced it
                           // (K6, or any host predictor); NULL: built from the pedestrian modes
};

__global__ void __launch_bounds__(64) pack_kernel(const Lay L, const SimArgs A, double* __restrict__ Pout)
{
    const int e = blockIdx.x;
    if (e >= A.n) return;
    const int tid = threadIdx.x, nt = blockDim.x;
    const int N = L.N;
    double* p = Pout + (size_t)e * L.np;
    const double sx = A.state[3 * e], sy = A.state[3 * e + 1], sth = A.state[3 * e + 2];
    __shared__ int s_idx;
    __shared__ int s_slot[64];
    if (tid == 0) {
        // closest reference sample within [idx - N, idx + 5N): the reference calls get_ref_states
        // with N_hor in the `action_steps` slot (trajectory_tracker.py:187,253-254)
        const int len = A.ref_len[e];
        const int idx0 = A.idx_ref[e];
        const int lb = idx0 - N > 0 ? idx0 - N : 0;
        const int ub = len < idx0 + 5 * N ? len : idx0 + 5 * N;
        double best = INFINITY;
        int bi = lb;
        for (int i = lb; i < ub; ++i) {
            const double* r = A.ref_traj + ((size_t)e * A.T + i) * 3;
            const double dx = sx - r[0], dy = sy - r[1];
            const double d = sqrt(dx * dx + dy * dy);
            if (d < best) { best = d; bi = i; }
        }
        s_idx = bi;
        A.idx_ref[e] = bi;
    }
    // rank the polygons by their distance to the robot, measured to the EDGES as
    // utils_geo.lineseg_dists does for MpcInterface.get_closest_n_stc_obstacles (utils_geo.py:6-33,
    // mpc_interface.py:90-100); stable: ties by index
    const int np_ = A.n_poly[e];
    for (int i = tid; i < 64; i += nt) s_slot[i] = -1;
    __syncthreads();
    __shared__ double s_pd[64];
    for (int i = tid; i < np_ && i < 64; i += nt) {
        const double* v = A.polys + ((size_t)e * A.Kp + i) * 8;
        double m = INFINITY;
        for (int k = 0; k < 4; ++k) {
            const int k1 = (k + 1) & 3;
            const double ax = v[2 * k], ay = v[2 * k + 1], bx = v[2 * k1], by = v[2 * k1 + 1];
            const double ex = bx - ax, ey = by - ay;
            const double ln = sqrt(ex * ex + ey * ey);
            const double dx = ex / ln, dy = ey / ln;
            const double s_ = (ax - sx) * dx + (ay - sy) * dy;
            const double t_ = (sx - bx) * dx + (sy - by) * dy;
            double h = s_ > t_ ? s_ : t_;
            h = h > 0.0 ? h : 0.0;
            const double c = (sx - ax) * dy - (sy - ay) * dx;
            const double d = sqrt(h * h + c * c);
            m = d < m ? d : m;
        }
        s_pd[i] = m;
    }
    __syncthreads();
    for (int i = tid; i < np_ && i < 64; i += nt) {
        int rank = 0;
        for (int j = 0; j < np_ && j < 64; ++j)
            rank += (s_pd[j] < s_pd[i]) || (s_pd[j] == s_pd[i] && j < i);
        if (rank < L.Nstc) s_slot[rank] = i;
    }
    __syncthreads();
    const int idx = s_idx;
    const int len = A.ref_len[e];
    // u_m1, s_0, s_N, q
    if (tid == 0) {
        p[L.p_um1] = A.last_u[2 * e]; p[L.p_um1 + 1] = A.last_u[2 * e + 1];
        p[L.p_s0] = sx; p[L.p_s0 + 1] = sy; p[L.p_s0 + 2] = sth;
        const int il = idx + N - 1 < len ? idx + N - 1 : len - 1;
        const double* r = A.ref_traj + ((size_t)e * A.T + il) * 3;
        p[L.p_sN] = r[0]; p[L.p_sN + 1] = r[1]; p[L.p_sN + 2] = r[2];
        for (int i = 0; i < 10; ++i) p[L.p_q + i] = A.tuning[i];
    }
    // r_s window (padded with the last sample), r_v, weights
    {
        const double gx = sx - A.goal[2 * e], gy = sy - A.goal[2 * e + 1];
        const double dist_to_goal = sqrt(gx * gx + gy * gy);
        double speed_ref;
        if (dist_to_goal >= A.base_speed * N * A.ts) speed_ref = A.base_speed;
        else {
            const double sr = dist_to_goal / N / A.ts;
            speed_ref = sr > A.lin_vel_max ? sr : A.lin_vel_max;     // the reference's max()
        }
        for (int k = tid; k < N; k += nt) {
            const int i = idx + k < len ? idx + k : len - 1;
            const double* r = A.ref_traj + ((size_t)e * A.T + i) * 3;
            p[L.p_rs + 3 * k] = r[0]; p[L.p_rs + 3 * k + 1] = r[1]; p[L.p_rs + 3 * k + 2] = r[2];
            p[L.p_rv + k] = speed_ref;
            p[L.p_qstc + k] = A.stc_w;
            p[L.p_qdyn + k] = A.dyn_w;
        }
    }
    // other robots: the reference's default zeros (trajectory_tracker.py:295-296)
    for (int i = tid; i < 3 * L.Nother * (N + 1); i += nt) p[L.p_c0 + i] = 0.0;
    // o_s: selected polygons as half-spaces b, a0, a1 (1 at the centroid, 0 on the edge)
    for (int i = tid; i < L.Nstc; i += nt) {
        double* o = p + L.p_os + i * 3 * L.nedge;
        const int src = s_slot[i];
        if (src < 0 || L.nedge != 4) {
            for (int k = 0; k < 3 * L.nedge; ++k) o[k] = 0.0;
            continue;
        }
        const double* v = A.polys + ((size_t)e * A.Kp + src) * 8;
        const double cx = (((v[0] + v[2]) + v[4]) + v[6]) / 4.0;
        const double cy = (((v[1] + v[3]) + v[5]) + v[7]) / 4.0;
        for (int k = 0; k < 4; ++k) {
            const int k1 = (k + 1) & 3;
            const double p0x = v[2 * k] - cx, p0y = v[2 * k + 1] - cy;
            const double p1x = v[2 * k1] - cx, p1y = v[2 * k1 + 1] - cy;
            const double det = p0x * p1y - p0y * p1x;
            const double ax = (p1y - p0y) / det, ay = (p0x - p1x) / det;
            o[k] = ax * cx + ay * cy + 1.0;
            o[4 + k] = ax;
            o[8 + k] = ay;
        }
    }
    // o_d: obstacle = (pedestrian, mode); slot t = position + (t*ts)*v, radius size + 0.03 t
    const int nobs = A.Pd * A.M < L.Ndyn ? A.Pd * A.M : L.Ndyn;
    for (int i = tid; i < L.Ndyn * (N + 1); i += nt) {
        const int ob = i / (N + 1), t = i - ob * (N + 1);
        double* o = p + L.p_od + (size_t)i * 6;
        if (A.od_in) {                       // supplied by the predictor stage
            const double* src = A.od_in + ((size_t)e * L.Ndyn * (N + 1) + i) * 6;
            for (int k = 0; k < 6; ++k) o[k] = src[k];
            continue;
        }
        if (ob >= nobs) { for (int k = 0; k < 6; ++k) o[k] = 0.0; continue; }
        const int pd = ob / A.M, md = ob - pd * A.M;
        const double* pos = A.ped_pos + ((size_t)e * A.Pd + pd) * 2;
        const double* vel = A.ped_vel + (((size_t)e * A.Pd + pd) * A.M + md) * 2;
        if (t == 0) {
            o[0] = pos[0]; o[1] = pos[1]; o[2] = A.ped_size; o[3] = A.ped_size;
        } else {
            const double tt = (double)t * A.ts;
            const double r = A.ped_size + 0.03 * (double)t;
            o[0] = pos[0] + tt * vel[0]; o[1] = pos[1] + tt * vel[1]; o[2] = r; o[3] = r;
        }
        o[4] = 0.0; o[5] = 1.0;
    }
}

// one thread per episode; u is the solver output [n, 2N] (one start per episode)
__global__ void plant_kernel(const SimArgs A, int N2, const double* __restrict__ u)
{
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= A.n || A.done[e]) return;
    const double v_sol = u[(size_t)e * N2], w_sol = u[(size_t)e * N2 + 1];
    // the sim loop never drives backwards: a negative speed is replaced by a full stop
    // (main_base.py:320-321); the tracker still remembers the solver's own action (:329)
    const double v = v_sol < 0.0 ? 0.0 : v_sol, w = v_sol < 0.0 ? 0.0 : w_sol;
    double s[3] = {A.state[3 * e], A.state[3 * e + 1], A.state[3 * e + 2]};
    const double ts = A.ts;
    double k1[3], k2[3], k3[3], k4[3], sn, cs;
    sincos_cw(s[2], &sn, &cs);
    k1[0] = ts * (v * cs); k1[1] = ts * (v * sn); k1[2] = ts * w;
    sincos_cw(s[2] + 0.5 * k1[2], &sn, &cs);
    k2[0] = ts * (v * cs); k2[1] = ts * (v * sn); k2[2] = ts * w;
    sincos_cw(s[2] + 0.5 * k2[2], &sn, &cs);
    k3[0] = ts * (v * cs); k3[1] = ts * (v * sn); k3[2] = ts * w;
    sincos_cw(s[2] + k3[2], &sn, &cs);
    k4[0] = ts * (v * cs); k4[1] = ts * (v * sn); k4[2] = ts * w;
#pragma unroll
    for (int i = 0; i < 3; ++i) s[i] = s[i] + (1.0 / 6.0) * (k1[i] + 2.0 * k2[i] + 2.0 * k3[i] + k4[i]);
    A.state[3 * e] = s[0]; A.state[3 * e + 1] = s[1]; A.state[3 * e + 2] = s[2];
    A.last_u[2 * e] = v_sol; A.last_u[2 * e + 1] = w_sol;
    for (int pd = 0; pd < A.Pd; ++pd) {
        double* pos = A.ped_pos + ((size_t)e * A.Pd + pd) * 2;
        const double* vel = A.ped_vel + ((size_t)e * A.Pd + pd) * A.M * 2;   // mode 0 is what happens
        pos[0] = pos[0] + ts * vel[0];
        pos[1] = pos[1] + ts * vel[1];
    }
    if (fabs(s[0] - A.goal[2 * e]) <= 0.5 && fabs(s[1] - A.goal[2 * e + 1]) <= 0.5 && fabs(v) < 0.4) A.done[e] = 1;
}

// K6 cluster_kernel: SWTA position hypotheses -> dynamic-obstacle slots.
// One thread per (episode, time offset): DBSCAN(eps, min_samples) exactly as
// sklearn.cluster.DBSCAN labels (utils_test.py:133-143; clusters numbered by their first core
// point), then per cluster mean and enlarge*std (utils_test.py:145-151, numpy's row-order sums)
// written as [mx, my, sx, sy, 0, 1] into o_d[cluster][t] (main_base.py:293-302).  Offsets with
// fewer clusters leave [0,0,0,0,0,1] in the slots of used obstacles, unused obstacles stay 0.
constexpr int CL_MAXK = 64;

__global__ void cluster_kernel(int n, int N, int K, int H, int Ndyn, double eps, int min_samples,
                               double enlarge, double human_size, const double* __restrict__ hyp,
                               const int* __restrict__ n_hyp, const double* __restrict__ cur,
                               double* __restrict__ od, int* __restrict__ ncl_out)
{
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n * (N + 1)) return;
    const int e = g / (N + 1), t = g - e * (N + 1);
    double* out = od + (size_t)e * Ndyn * (N + 1) * 6;
    if (t == 0) {                       // current positions with the pedestrians' size
        for (int h = 0; h < H && h < Ndyn; ++h) {
            double* o = out + ((size_t)h * (N + 1)) * 6;
            o[0] = cur[((size_t)e * H + h) * 2]; o[1] = cur[((size_t)e * H + h) * 2 + 1];
            o[2] = human_size; o[3] = human_size; o[4] = 0.0; o[5] = 1.0;
        }
        ncl_out[(size_t)e * (N + 1)] = H < Ndyn ? H : Ndyn;
        return;
    }
    const double* X = hyp + ((size_t)(e * N + (t - 1)) * K) * 2;
    const int k = n_hyp ? n_hyp[e * N + (t - 1)] : K;
    signed char label[CL_MAXK];
    unsigned long long nb[CL_MAXK];
    unsigned char stack[CL_MAXK];
    const double eps2 = eps * eps;
    for (int i = 0; i < k; ++i) {
        unsigned long long m = 0ull;
        for (int j = 0; j < k; ++j) {
            const double dx = X[2 * i] - X[2 * j], dy = X[2 * i + 1] - X[2 * j + 1];
            if (dx * dx + dy * dy <= eps2) m |= 1ull << j;
        }
        nb[i] = m;
        label[i] = -1;
    }
    int c = 0;
    for (int i = 0; i < k; ++i) {
        if (label[i] != -1 || __popcll(nb[i]) < min_samples) continue;
        label[i] = (signed char)c;
        int sp = 0;
        stack[sp++] = (unsigned char)i;
        while (sp) {
            const int j = stack[--sp];
            if (__popcll(nb[j]) >= min_samples) {
                unsigned long long m = nb[j];
                while (m) {
                    const int v = __ffsll((long long)m) - 1;
                    m &= m - 1;
                    if (label[v] == -1) { label[v] = (signed char)c; stack[sp++] = (unsigned char)v; }
                }
            }
        }
        ++c;
    }
    ncl_out[(size_t)e * (N + 1) + t] = c < Ndyn ? c : Ndyn;
    for (int q = 0; q < c && q < Ndyn; ++q) {
        double m0 = 0.0, m1 = 0.0;
        int cnt = 0;
        for (int i = 0; i < k; ++i)
            if (label[i] == q) { m0 += X[2 * i]; m1 += X[2 * i + 1]; ++cnt; }
        m0 /= cnt; m1 /= cnt;
        double v0 = 0.0, v1 = 0.0;
        for (int i = 0; i < k; ++i)
            if (label[i] == q) {
                v0 += (X[2 * i] - m0) * (X[2 * i] - m0);
                v1 += (X[2 * i + 1] - m1) * (X[2 * i + 1] - m1);
            }
        double* o = out + ((size_t)q * (N + 1) + t) * 6;
        o[0] = m0; o[1] = m1;
        o[2] = sqrt(v0 / cnt) * enlarge + 0.0;
        o[3] = sqrt(v1 / cnt) * enlarge + 0.0;
        o[4] = 0.0; o[5] = 1.0;
    }
}

// second pass: slots of used obstacles that no cluster filled become [0,0,0,0,0,1]
__global__ void cluster_fill_kernel(int n, int N, int Ndyn, const int* __restrict__ ncl, double* __restrict__ od)
{
    const int e = blockIdx.x;
    if (e >= n) return;
    int nobs = 0;
    for (int t = 0; t <= N; ++t) nobs = ncl[(size_t)e * (N + 1) + t] > nobs ? ncl[(size_t)e * (N + 1) + t] : nobs;
    for (int i = threadIdx.x; i < Ndyn * (N + 1); i += blockDim.x) {
        const int ob = i / (N + 1), t = i - ob * (N + 1);
        double* o = od + ((size_t)e * Ndyn * (N + 1) + i) * 6;
        if (ob >= nobs) { for (int k = 0; k < 6; ++k) o[k] = 0.0; }
        else if (ob >= ncl[(size_t)e * (N + 1) + t]) { o[0] = o[1] = o[2] = o[3] = o[4] = 0.0; o[5] = 1.0; }
    }
}

}  // namespace mpcb
