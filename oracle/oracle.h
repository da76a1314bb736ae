/*
 * oracle.h -- C interface of the CPU ORACLE for the uneven_planner back-end hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product (uneven_planner_b200/) may include, link
 * or call this.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs use it, as the checker / the timed CPU baseline.
 *
 * The oracle is a plain-C++ (no Eigen, no ROS) fp64 restatement of the reference algorithm:
 *   ALMTrajOpt::optimizeSE2Traj      src/uneven_planner/back_end/src/alm_traj_opt.cpp:168-278
 *   innerCallback                    alm_traj_opt.cpp:280-347
 *   ALMTrajOpt::initScaling          alm_traj_opt.cpp:349-661
 *   ALMTrajOpt::calConstrainCostGrad alm_traj_opt.cpp:663-991
 *   updateDualVars/judgeConvergence  back_end/include/back_end/alm_traj_opt.h:132-163
 *   expC2/logC2/getTtoTauGrad        alm_traj_opt.h:232-261
 *   MinJerkOpt / MINCO_SE2           back_end/include/utils/se2traj.hpp:564-870
 *   BandedSystem                     back_end/include/utils/banded_system.hpp:25-145
 *   lbfgs_optimize + line search     back_end/include/utils/lbfgs.hpp:276-722
 *   UnevenMap::getAllWithGrad chain  uneven_map/include/uneven_map/uneven_map.h:258-377,398-454
 *
 * PARITY PINNED AGAINST THE REFERENCE'S OWN SOURCE (round 1): the reference ships no tests, golden vectors or fixtures for this
 * path and its dependencies (Eigen, ROS, PCL, OMPL) are absent from this image, but its back-end sources
 * compile UNMODIFIED against oracle/shim (a minimal Eigen stand-in with this oracle's reduction orders, no-op ROS/PCL headers):
 * `make -C oracle ref` builds back_end/src/alm_traj_opt.cpp + back_end/include/{back_end/alm_traj_opt.h, utils/*.hpp} +
 * uneven_map/include/uneven_map/uneven_map.h from /root/reference into oracle/_ref/libref.so, and tests/test_ref_pin.py requires
 * bit-identical outputs from that build and from this restatement: MinJerkOpt / BandedSystem / lbfgs_optimize piecewise, and
 * whole ALMTrajOpt::optimizeSE2Traj solves (return code, coefficients, durations, multipliers, scales).  What that build does
 * NOT contain: real Eigen (its SIMD reduction order is version- and flag-dependent; the shim fixes the canonical order below),
 * libm trig (replaced by include/ualm_detmath.h on all three sides), uneven_map.cpp (needs PCL: the grid is injected and the
 * six-line normSO2 is restated in oracle/ref_alm_driver.cpp).  Analytic invariants and finite-difference checks
 * (tests/test_oracle_pins.py) pin the mathematics independently.
 */
#ifndef UALM_ORACLE_H
#define UALM_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

/* alm_traj_opt/ * parameters (alm_traj_opt.cpp:7-27) + gravity (uneven_map.cpp:85). */
typedef struct {
    double rho_T, rho_ter, max_vel, max_acc_lon, max_acc_lat, max_kap, min_cxi, max_sig;
    int    use_scaling;
    double rho, beta, gamma, epsilon_con, max_iter;
    double g_epsilon, min_step, inner_max_iter, delta;
    int    mem_size, past, int_K;
    double gravity;
} orc_params_t;

/* UnevenMap geometry (uneven_map.cpp:96-114) + grid; cells = [X][Y][Yaw][4] {z, sigma, zbx, zby}
 * in the reference address order x*Y*Yaw + y*Yaw + yaw (uneven_map.h:427-435). */
typedef struct {
    const double *cells;
    int    voxel_num[3];
    double origin[3];        /* = min_boundary */
    double max_boundary[3];
    double xy_resolution, yaw_resolution;
} orc_map_t;

/* One optimizeSE2Traj problem (alm_traj_opt.h:92-98).  Matrices are column-major like Eigen:
 * init_xy/end_xy = 2x3 [px,py, vx,vy, ax,ay]; inner_xy = 2x(N-1) [x1,y1,x2,y2,...]. */
typedef struct {
    int    N, M;             /* piece_xy, piece_yaw */
    const double *init_xy, *end_xy, *inner_xy;
    const double *init_yaw, *end_yaw, *inner_yaw;
    double total_time;
} orc_problem_t;

typedef struct {
    int    ret_code;         /* 0 ok / 1 L-BFGS hard error / 2 ALM max-iter (alm_traj_opt.cpp:176,252,267) */
    int    outer_iters;      /* `iter` at exit */
    int    n_evals;          /* innerCallback calls (initScaling not counted) */
    int    n_lbfgs_iters;    /* accepted line searches over all inner solves */
    int    last_lbfgs_ret;
    int    max_bound;        /* largest L-BFGS history depth reached */
    double inner_cost;       /* f returned by the last lbfgs_optimize */
    double jerk_cost;        /* minco_se2.getTrajJerkCost() at exit (state of LAST evaluation) */
    double total_T;          /* sum of piece durations at exit */
    double res_h, res_g;     /* the two norms of judgeConvergence at exit */
    double scale_fx;
    double rho_final;
    double t_total, t_minco, t_penalty, t_adjoint, t_lbfgs, t_scaling; /* seconds (own timers) */
} orc_result_t;

/* Full solve.  c_xy_out: 6N x 2 column-major (x coeffs then y coeffs, low->high power per piece,
 * se2traj.hpp:585); c_yaw_out: 6M; x_out: 1+2(N-1)+(M-1) decision vector [tau | Pxy | Pyaw].
 * Optional (may be NULL): lambda_out[S], mu_out[6S], scale_cx_out[7S]. */
int orc_solve(const orc_params_t *p, const orc_map_t *map, const orc_problem_t *prob,
              orc_result_t *res, double *c_xy_out, double *c_yaw_out, double *x_out,
              double *lambda_out, double *mu_out, double *scale_cx_out);

/* Same in float arithmetic (experimental; NOT the oracle of record). */
int orc_solve_f32(const orc_params_t *p, const orc_map_t *map, const orc_problem_t *prob,
                  orc_result_t *res, double *c_xy_out, double *c_yaw_out, double *x_out);

/* One innerCallback evaluation at x with given duals/scales (kernel-level parity).
 * lambda[S], mu[6S], scale_cx[7S] in reference layout; outputs: f, grad[n], hx[S], gx[6S],
 * parts[3] = {jerk_cost, constrain_cost, tau_cost}, c_xy, c_yaw, and the (C,T)-gradients of the
 * constraint term before the adjoint: gdCxy[12N], gdTxy[N], gdCyaw[6M], gdTyaw[M] (may be NULL). */
int orc_eval(const orc_params_t *p, const orc_map_t *map, const orc_problem_t *prob,
             const double *x, const double *lambda, const double *mu, const double *scale_cx,
             double scale_fx, double rho,
             double *f, double *grad, double *hx, double *gx, double *parts,
             double *c_xy_out, double *c_yaw_out,
             double *gdCxy, double *gdTxy, double *gdCyaw, double *gdTyaw);

/* initScaling at x0 (alm_traj_opt.cpp:349-661): scale_fx and scale_cx[7S]. */
int orc_init_scaling(const orc_params_t *p, const orc_map_t *map, const orc_problem_t *prob,
                     const double *x0, double *scale_fx, double *scale_cx);

/* UnevenMap::getAllWithGrad (uneven_map.h:318-377): values[7], grads[7*3] (row-major 7x3). */
void orc_map_query(const orc_map_t *map, const double pos[3], double *values, double *grads);

/* MINCO pieces (se2traj.hpp:595-816) for unit tests: Dim in {1,2}, general (non-uniform) ts[N].
 * inPs: Dim x (N-1) col-major; head/tail: Dim x 3 col-major; c: 6N x Dim col-major. */
int orc_minco_generate(int Dim, int N, const double *inPs, const double *ts,
                       const double *head, const double *tail, double *c);
double orc_minco_jerk(int Dim, int N, const double *c, const double *ts, double *gdC, double *gdT);
/* gdC (6N x Dim col-major) and gdT[N] (in/out) -> gdP (Dim x (N-1) col-major). */
int orc_minco_grad_ct_to_qt(int Dim, int N, const double *inPs, const double *ts,
                            const double *head, const double *tail,
                            const double *gdC, double *gdT, double *gdP);

/* lbfgs_optimize on the n-dim Rosenbrock function (test pin 6). returns lbfgs code. */
int orc_lbfgs_rosenbrock(int n, double *x, double *f, int mem_size, double g_epsilon, int past,
                         double delta, int *iters);

/* Post-solve quality scan: ALMTrajOpt::getMaxVxAxAyCurAttSig (alm_traj_opt.h:170-229) and
 * SE2Trajectory::getNonHolError (se2traj.hpp:551-561) of the trajectory {c_xy (6N x 2 col-major), c_yaw (6M), uniform piece
 * durations T_xy, T_yaw}, sampled every dt (0.01 in the reference).
 * out[8] = {max_vx, max_ax, max_ay, max_cur, max_att, max_sig, nonhol_error, number of samples}. */
void orc_feasibility(const orc_map_t *map, double gravity, int N, int M, const double *c_xy, const double *c_yaw,
                     double T_xy, double T_yaw, double dt, double *out);

double orc_expC2(double tau);
double orc_logC2(double T);
double orc_dTdtau(double tau);

#ifdef __cplusplus
}
#endif
#endif
