/*
 * mpcb.h — C-ABI of the B200-native batched NMPC (PANOC + ALM/PM) solver.
 *
 * This is the drop-in boundary for the ONE hot path of
 * Woodenonez/DyObAv-MPCnWTA-Warehouse: the OpEn-generated solver that
 * `src/pkg_mpc_tracker/trajectory_tracker.py:61-62` loads and `:362` calls
 * (`self.solver.run(parameters)`).  The reference's generated extension is a
 * PyO3 module wrapping `solve(p, cache, u, y0, c0)`; these entry points are
 * what a binding for that call would bind, batched over B independent
 * instances.  Plain pointers and sizes only — no torch types.
 *
 * All `const double*` / `double*` / `int32_t*` arguments of the *_f64 entry
 * points are DEVICE pointers (e.g. torch.Tensor.data_ptr()) unless the name
 * ends in `_host`.  Every function returns 0 on success or a negative
 * MPCB_E_* code; nothing throws, nothing allocates behind the caller's back
 * (the caller provides `workspace`) except the *_host convenience path, and all device work is ordered on the
 * `stream` argument (a cudaStream_t passed as void*).
 */
#ifndef MPCB_H
#define MPCB_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MPCB_ABI_VERSION 4

/* ---- error codes --------------------------------------------------------- */
#define MPCB_OK            0
#define MPCB_E_DIMS       (-1)  /* unsupported / inconsistent dimensions        */
#define MPCB_E_NULL       (-2)  /* required pointer is NULL                     */
#define MPCB_E_WORKSPACE  (-3)  /* workspace too small                          */
#define MPCB_E_CUDA       (-4)  /* CUDA runtime error (see mpcb_last_error)     */
#define MPCB_E_ALIGN      (-5)  /* pointer not 8-byte aligned                   */
#define MPCB_E_NO_DEVICE  (-6)  /* no CUDA device / wrong architecture          */

/* ---- per-instance exit status (mirrors OpEn's ExitStatus, the strings that
 *      config/mpc_fast.yaml:48 `bad_exit_codes` refers to) ------------------- */
#define MPCB_CONVERGED                   0
#define MPCB_NOT_CONVERGED_ITERATIONS    1
#define MPCB_NOT_CONVERGED_OUT_OF_TIME   2  /* only with cfg->max_inner_total > 0 (iteration budget) or cfg->max_time_us > 0 */
#define MPCB_NOT_FINITE_COMPUTATION      3

/*
 * Problem dimensions.  The reference fixes these at code-generation time from
 * config/mpc_*.yaml (`N_hor, Nother, Nstcobs, nstcobs/3, Ndynobs`, lines
 * 21-31); here they are run-time.  nu=2, ns=3, nq=10, ndynobs=6 are the
 * unicycle problem's constants (mpc_builder.py:45-58).
 */
typedef struct mpcb_dims {
    int32_t N;       /* horizon N_hor (1..64)                                   */
    int32_t Nother;  /* other robots                                            */
    int32_t Nstc;    /* static polygons (half-space form)                       */
    int32_t nedge;   /* edges per polygon = nstcobs/3 (reference: 4), 1..8      */
    int32_t Ndyn;    /* dynamic-obstacle ellipses per time offset (<= 256)      */
} mpcb_dims;

/* Robot / kinematic constants: config/mpc_fast.yaml:6-18 (configs.py:93-103). */
typedef struct mpcb_robot {
    double ts;
    double vehicle_width;   /* fleet safe distance (mpc_builder.py:90,97)       */
    double vehicle_margin;  /* ellipse inflation   (mpc_builder.py:122,140)     */
    double social_margin;   /* extra inflation for the t=0 ellipse (:122)       */
    double lin_vel_min, lin_vel_max;   /* input box U (:151-153)                */
    double ang_vel_max;
    double lin_acc_min, lin_acc_max;   /* ALM set C  (:162-166)                 */
    double ang_acc_max;
} mpcb_robot;

/*
 * Solver settings: OpEn SolverConfiguration as the reference leaves it
 * (mpc_builder.py:187-195): only initial_penalty=10 and the wall-clock cap are
 * set, the rest are opengen 0.6.13 defaults.
 */
typedef struct mpcb_solver_cfg {
    double tolerance;            /* eps           1e-4                          */
    double initial_tolerance;    /* eps_0         1e-4                          */
    double delta_tolerance;      /* delta         1e-4                          */
    double inner_tol_update;     /* beta          0.1                           */
    double penalty_update;       /* rho           5.0                           */
    double sufficient_decrease;  /* theta         0.1                           */
    double initial_penalty;      /* c0            10.0 (mpc_builder.py:188)     */
    double sy_epsilon;           /* L-BFGS s'y    1e-10                         */
    double cbfgs_epsilon;        /* C-BFGS eps    1e-8                          */
    double cbfgs_alpha;          /* C-BFGS alpha  1.0                           */
    int32_t max_inner;           /* 500                                         */
    int32_t max_outer;           /* 10                                          */
    int32_t lbfgs_mem;           /* 10 (<= MPCB_MAX_LBFGS)                      */
    int32_t max_inner_total;     /* 0 = off.  Budget on the inner (PANOC) iterations of ONE solve,
                                  * summed over the outer iterations: the batch analogue of the
                                  * reference's wall-clock cap max_solver_time (mpc_fast.yaml:45,
                                  * mpc_builder.py:189).  The clock is the iteration count: the inner
                                  * loop stops when the budget is used up (inner status
                                  * NotConvergedOutOfTime), the ALM step finishes as usual and the
                                  * next outer iteration is refused with MPCB_NOT_CONVERGED_OUT_OF_TIME,
                                  * the status mpc_fast.yaml:48 lists in bad_exit_codes.             */
    int32_t team_mode;           /* 0 = by the dimensions (team kernels from 64 ellipses on), 1 = team
                                  * kernels for ANY dimensions - one instance gets a whole CTA (a solver
                                  * warp plus the worker pool) instead of one warp.  Selects the team
                                  * arithmetic contract (mpcb_team_groups_cfg), so results differ from
                                  * team_mode 0 by round-off.  Meant as a latency mode for single solves;
                                  * measured on B200 it pays only when the per-step part of an evaluation
                                  * is large (reference dims: 28.7 vs 20.8 us per inner iteration).     */
    int32_t max_time_us;         /* 0 = off.  Wall-clock cap on ONE solve in microseconds, the reference's
                                  * max_solver_time itself (with_max_duration_micros, mpc_builder.py:189;
                                  * 100 000 in mpc_fast.yaml:45): the time since the instance was started is
                                  * checked before every outer iteration and after every inner iteration, as
                                  * AlmOptimizer::solve / PANOCOptimizer::solve do; exhausted ->
                                  * MPCB_NOT_CONVERGED_OUT_OF_TIME.  Results then depend on timing (as the
                                  * reference's do).  Honoured by the latency kernel, i.e. for batches of up to
                                  * twelve instances per SM and dims without team kernels - the single solve per
                                  * timestep; larger batches ignore it (max_inner_total is their budget).   */
} mpcb_solver_cfg;

#define MPCB_MAX_LBFGS 10
#define MPCB_MAX_N     64
#define MPCB_MAX_EDGE  8
#define MPCB_MAX_NDYN  256

/* Length of the parameter vector p for these dims (2778 at the yaml defaults). */
int32_t mpcb_param_len(const mpcb_dims* dims);
/* nu*N, n1 (=2N ALM constraints), n2 (=max(Ndyn,1) penalty constraints). */
int32_t mpcb_num_decision(const mpcb_dims* dims);
int32_t mpcb_n1(const mpcb_dims* dims);
int32_t mpcb_n2(const mpcb_dims* dims);

/* Worker groups G of the team kernels for these dims (0: the one-warp kernels are used).  Part of
 * the arithmetic contract: with G > 0 the ellipse cost terms of a horizon step are summed per
 * group i % G first (csrc/mpcb_device.cuh "team mode"); the laned oracle mirrors the rule. */
int32_t mpcb_team_groups(const mpcb_dims* dims);
/* ... and with cfg->team_mode taken into account (what mpcb_solve / mpcb_eval will really use). */
int32_t mpcb_team_groups_cfg(const mpcb_dims* dims, const mpcb_solver_cfg* cfg);

int32_t mpcb_abi_version(void);
/* Text of the last CUDA error seen by this thread ("" if none). */
const char* mpcb_last_error(void);
/* Fill defaults equal to the reference's yaml / opengen defaults. */
void mpcb_default_robot(mpcb_robot* r);
void mpcb_default_solver_cfg(mpcb_solver_cfg* c);

/*
 * Device workspace (bytes) the eval/solve entry points need for n_p parameter
 * rows: the staged structure-of-arrays copy of every row (K3 writes it, the
 * solve kernel reads it through L1 or TMA-loads it) behind a header:
 *   [0, MPCB_WS_COUNTER_BYTES)        work-queue counters (int32, one per persistent CTA)
 *   then uint64[MPCB_WS_PROF_CTAS]    launch profile of the last mpcb_solve_f64: the
 *                                     %globaltimer (ns) at which CTA c started, and
 *   uint64[MPCB_WS_PROF_CTAS][MPCB_WS_PROF_WARPS]  at which warp w of CTA c ran out of work
 *                                     (0: the warp did not take part) - read back by bench.py
 *                                     to report the launch tail
 * followed by the scenario order of the solve (int32[n_p], float[n_p] difficulty keys: the solve
 * kernel starts with the scenarios whose reference path an obstacle blocks, which are the ones
 * that use up the iteration caps) and then the staged blocks.
 */
#define MPCB_WS_COUNTER_BYTES 4096
#define MPCB_WS_PROF_CTAS     1024
#define MPCB_WS_PROF_WARPS    16
#define MPCB_WS_HEADER_BYTES  (MPCB_WS_COUNTER_BYTES + MPCB_WS_PROF_CTAS * (1 + MPCB_WS_PROF_WARPS) * 8)
int32_t mpcb_workspace_bytes(const mpcb_dims* dims, int32_t n_p, int32_t starts,
                             size_t* bytes);

/*
 * Evaluate the augmented cost the inner solver minimises and its pieces, for
 * B instances:  psi(u; c, y) = f(u) + c/2 [ dist^2_C(F1(u)+y/max(c,1)) + |F2(u)|^2 ].
 * Replaces the CasADi-generated `cost`, `grad`, `mapping_f1`, `mapping_f2`
 * C functions the OpEn build emits from mpc_builder.py:171-174.
 *   p [n_p, np]  reference layout (mpc_builder.py:60), one row per scenario
 *   u [B, 2N]    B = n_p * starts; instance b uses row b / starts of p
 *   y [B, n1] or NULL (zeros), c [B] or NULL (initial_penalty)
 * Outputs (any may be NULL): f[B], psi[B], grad[B,2N], F1[B,n1], F2[B,n2].
 */
int32_t mpcb_eval_f64(const mpcb_dims* dims, const mpcb_robot* robot,
                      const mpcb_solver_cfg* cfg,
                      int32_t n_p, int32_t starts,
                      const double* p, const double* u,
                      const double* y, const double* c,
                      double* f, double* psi, double* grad,
                      double* F1, double* F2,
                      void* workspace, size_t workspace_bytes,
                      void* stream);

/*
 * Solve B = n_p*starts independent instances: ALM/PM outer loop around PANOC,
 * replacing the generated `Solver.run(p, initial_guess,
 * initial_lagrange_multipliers, initial_penalty)` (trajectory_tracker.py:13-15).
 *   u0 [B,2N] or NULL (zeros, the reference's behaviour: it passes p only)
 *   y0 [B,n1] or NULL (zeros);  c0 [B] or NULL (cfg->initial_penalty)
 * Outputs (u_out, exit_status required; others may be NULL):
 *   u_out[B,2N] solution, cost[B] (= f(u*)), exit_status[B] (MPCB_* above),
 *   n_outer[B], n_inner[B], fpr[B] last inner |gamma*fpr|, f1_infeas[B]
 *   (=|y+ - y|/c), f2_norm[B], penalty[B] final c, y_out[B,n1] multipliers,
 *   evals[B,4]: {cost-only, cost+gradient} horizon evaluations the kernel performed
 *   (feeds the work/roofline accounting), the number of inner iterations that started
 *   with |gamma fpr| < tolerance but failed the AKKT test (so the solve went on), 0.
 * Kernel choice (results are bit-identical whichever runs, for a given dims/cfg): batches of at most twelve
 * instances per SM - the single solve per timestep of the reference, small fleets - run the latency kernel (a CTA
 * per instance, line-search trials evaluated concurrently by helper warps); larger batches the
 * one-warp-per-instance queue kernel; dims with >= 64 ellipses (or cfg->team_mode = 1) the team kernels.
 */
int32_t mpcb_solve_f64(const mpcb_dims* dims, const mpcb_robot* robot,
                       const mpcb_solver_cfg* cfg,
                       int32_t n_p, int32_t starts,
                       const double* p, const double* u0,
                       const double* y0, const double* c0,
                       double* u_out, double* cost, int32_t* exit_status,
                       int32_t* n_outer, int32_t* n_inner,
                       double* fpr, double* f1_infeas, double* f2_norm,
                       double* penalty, double* y_out, int32_t* evals,
                       void* workspace, size_t workspace_bytes,
                       void* stream);

/*
 * f32 twins: the same two entry points with float32 DEVICE buffers at the boundary.
 * The arithmetic stays f64 - PANOC's Lipschitz probe perturbs the iterate by 1e-12 and the 1e-6
 * radius regulariser scales the ellipse terms by up to 1e12, neither of
 * which single precision can represent - so this is a boundary mode: inputs are widened exactly on
 * the device, solved by the f64 kernels, results rounded once.  Stated tolerance: the outputs equal
 * mpcb_solve_f64's on the widened inputs, rounded to float32 (bit for bit); against an f64 solve of
 * the unrounded parameters the difference is the effect of rounding p to float32 (about 1e-6 m in
 * the positions), which PANOC carries to <= 1e-3 in u on converging instances (tests/test_f32.py).
 * Workspace: mpcb_workspace_bytes_f32 (+ 8*B*n2 bytes, rounded up to 256, for mpcb_eval_f32's F2),
 * 256-byte aligned.  exit_status / n_outer / n_inner / evals stay int32.
 */
int32_t mpcb_workspace_bytes_f32(const mpcb_dims* dims, int32_t n_p, int32_t starts, size_t* bytes);
int32_t mpcb_solve_f32(const mpcb_dims* dims, const mpcb_robot* robot, const mpcb_solver_cfg* cfg,
                       int32_t n_p, int32_t starts,
                       const float* p, const float* u0, const float* y0, const float* c0,
                       float* u_out, float* cost, int32_t* exit_status,
                       int32_t* n_outer, int32_t* n_inner,
                       float* fpr, float* f1_infeas, float* f2_norm,
                       float* penalty, float* y_out, int32_t* evals,
                       void* workspace, size_t workspace_bytes, void* stream);
int32_t mpcb_eval_f32(const mpcb_dims* dims, const mpcb_robot* robot, const mpcb_solver_cfg* cfg,
                      int32_t n_p, int32_t starts,
                      const float* p, const float* u, const float* y, const float* c,
                      float* f, float* psi, float* grad, float* F1, float* F2,
                      void* workspace, size_t workspace_bytes, void* stream);

/*
 * Host-buffer convenience for the single-solve path the reference actually
 * uses (one p list per timestep): copies p (and u0 if given) H2D, solves one
 * instance, copies the results back, synchronises.  All pointers are HOST.
 * out_scalars[8] = {cost, fpr, f1_infeas, f2_norm, penalty, n_outer, n_inner,
 * solve_time_ms (device time of the solve kernel)}.
 */
int32_t mpcb_solve_one_host(const mpcb_dims* dims, const mpcb_robot* robot,
                            const mpcb_solver_cfg* cfg,
                            const double* p_host, const double* u0_host,
                            const double* y0_host, const double* c0_host,
                            double* u_out_host, double* y_out_host,
                            int32_t* exit_status_host, double* out_scalars);

/* ---- closed-loop support: device-side packer and plant ----------
 * Episode state for E parallel receding-horizon episodes; every pointer is a DEVICE pointer.
 * mpcb_pack_f64 writes one reference-layout parameter row per episode, replacing the host
 * work of TrajectoryTracker.run_step (trajectory_tracker.py:285-317), get_ref_states
 * (:243-270) and MpcInterface.get_stc_constraints / get_dyn_constraints
 * (mpc_interface.py:73-100); mpcb_plant_step_f64 applies the first action with the RK4
 * unicycle (motion_model.py:141-163, basic_agent.py:106), advances the pedestrians and
 * evaluates the termination test (trajectory_tracker.py:191-199). */
typedef struct mpcb_sim {
    int32_t n;              /* episodes                                              */
    int32_t T;              /* padded length of every reference trajectory           */
    int32_t Kp;             /* polygon slots per episode (4 vertices each, <= 64)    */
    int32_t Pd, M;          /* pedestrians per episode, predicted modes each         */
    double base_speed, lin_vel_max, ped_size, stc_w, dyn_w, ts;
    double tuning[10];      /* q block (trajectory_tracker.py:138-139)               */
    double* state;          /* [n,3] in/out                                          */
    double* last_u;         /* [n,2] in/out                                          */
    const double* ref_traj; /* [n,T,3]                                               */
    const int32_t* ref_len; /* [n]                                                   */
    int32_t* idx_ref;       /* [n] in/out                                            */
    const double* goal;     /* [n,2]                                                 */
    const double* polys;    /* [n,Kp,4,2] inflated rectangles/quadrilaterals         */
    const int32_t* n_poly;  /* [n]                                                   */
    double* ped_pos;        /* [n,Pd,2] in/out                                       */
    const double* ped_vel;  /* [n,Pd,M,2]; mode 0 is the motion that happens         */
    int32_t* done;          /* [n] in/out                                            */
    const double* od_in;    /* [n,Ndyn,N+1,6] or NULL: the o_d block from a predictor
                             * stage (mpcb_cluster_f64 or the host); NULL: built from
                             * the pedestrians' constant-velocity modes               */
} mpcb_sim;

int32_t mpcb_pack_f64(const mpcb_dims* dims, const mpcb_sim* sim, double* p_out, void* stream);
int32_t mpcb_plant_step_f64(const mpcb_dims* dims, const mpcb_sim* sim, const double* u, void* stream);
/* the portable sin/cos the kernels use, callable on the host (bit-identical) */
void mpcb_sincos_host(double x, double* sn, double* cs);

/* SWTA position hypotheses -> the o_d block, on the device.  Replaces
 * MainBase.run_wta_prediction's clustering (main_base.py:196-207: DBSCAN(eps=1, min_samples=2)
 * per time offset, utils_test.py:133-143, and the (mean, 2*std) fit, utils_test.py:145-151)
 * and run_one_step's slot list (main_base.py:293-302).
 *   hyp [n, N, K, 2] hypotheses of offsets 1..N (K <= 64), n_hyp [n, N] valid counts or NULL,
 *   cur_pos [n, H, 2] current pedestrian positions, o_d [n, Ndyn, N+1, 6] output,
 *   scratch [n, N+1] int32.  All DEVICE pointers. */
int32_t mpcb_cluster_f64(const mpcb_dims* dims, int32_t n, int32_t K, int32_t H, double eps,
                         int32_t min_samples, double enlarge, double human_size, const double* hyp,
                         const int32_t* n_hyp, const double* cur_pos, double* o_d, int32_t* scratch,
                         void* stream);

/* Measured FP64 FMA throughput of the current device (TFLOP/s, 2 flop per FMA): the roofline
 * denominator bench.py reports for the solve kernel. */
int32_t mpcb_fp64_peak_tflops(double* tflops);

#ifdef __cplusplus
}
#endif
#endif /* MPCB_H */
