/* oracle/ref_driver.cpp -- drives the reference's OWN back-end utility headers, compiled unmodified from /root/reference:
 *     back_end/include/utils/se2traj.hpp      (Piece / PolyTrajectory / MinJerkOpt<Dim>: se2traj.hpp:31-816)
 *     back_end/include/utils/banded_system.hpp (BandedSystem: banded_system.hpp:25-145)
 *     back_end/include/utils/lbfgs.hpp         (lbfgs_optimize + line_search_lewisoverton: lbfgs.hpp:276-722)
 * against oracle/shim (a minimal Eigen stand-in, empty ros/ros.h, stubbed root_finder.hpp) because Eigen and ROS are absent
 * from this image.  Built by `make -C oracle ref` into oracle/_ref/libref.so when /root/reference exists; tests/test_ref_pin.py
 * compares its outputs with the oracle's restatement bit for bit.  TEST INFRASTRUCTURE ONLY. */
#include "utils/se2traj.hpp"
#include "utils/lbfgs.hpp"

#include <cstring>

using namespace uneven_planner;

namespace {
template <int Dim>
struct Runner {
    MinJerkOpt<Dim> opt;
    void generate(int N, const double *inPs, const double *ts, const double *head, const double *tail)
    {
        opt.reset(N);
        Eigen::MatrixXd P(Dim, N - 1), H(Dim, 3), T(Dim, 3);
        Eigen::VectorXd t(N);
        for (int j = 0; j < N - 1; j++) for (int d = 0; d < Dim; d++) P(d, j) = inPs[d + (size_t)j * Dim];
        for (int j = 0; j < 3; j++) for (int d = 0; d < Dim; d++) { H(d, j) = head[d + j * Dim]; T(d, j) = tail[d + j * Dim]; }
        for (int i = 0; i < N; i++) t(i) = ts[i];
        opt.generate(P, t, H, T);
    }
};
template <int Dim>
int minco_all(int N, const double *inPs, const double *ts, const double *head, const double *tail, double *c, double *jerk, double *gdC_jerk,
              double *gdT_jerk, const double *gdC_in, double *gdT_io, double *gdP)
{
    Runner<Dim> r;
    r.generate(N, inPs, ts, head, tail);
    const Eigen::MatrixXd &cc = r.opt.getCoeffs();
    for (int d = 0; d < Dim; d++) for (int i = 0; i < 6 * N; i++) c[i + (size_t)d * 6 * N] = cc(i, d);
    if (jerk) *jerk = r.opt.getTrajJerkCost();
    if (gdC_jerk) {
        Eigen::MatrixXd gC;
        Eigen::VectorXd gT;
        r.opt.calJerkGradCT(gC, gT);
        for (int d = 0; d < Dim; d++) for (int i = 0; i < 6 * N; i++) gdC_jerk[i + (size_t)d * 6 * N] = gC(i, d);
        for (int i = 0; i < N; i++) gdT_jerk[i] = gT(i);
    }
    if (gdC_in) {
        Eigen::MatrixXd gC(6 * N, Dim), gP;
        Eigen::VectorXd gT(N);
        for (int d = 0; d < Dim; d++) for (int i = 0; i < 6 * N; i++) gC(i, d) = gdC_in[i + (size_t)d * 6 * N];
        for (int i = 0; i < N; i++) gT(i) = gdT_io[i];
        r.opt.calGradCTtoQT(gC, gT, gP);
        for (int i = 0; i < N; i++) gdT_io[i] = gT(i);
        for (int j = 0; j < N - 1; j++) for (int d = 0; d < Dim; d++) gdP[d + (size_t)j * Dim] = gP(d, j);
    }
    return 0;
}

struct Rosen { int n; int evals; };
double rosen_eval(void *inst, const Eigen::VectorXd &v, Eigen::VectorXd &g)
{   /* same arithmetic as orc_lbfgs_rosenbrock (oracle.cpp) */
    Rosen *r = (Rosen *)inst;
    r->evals++;
    double fx = 0.0;
    for (int i = 0; i < r->n; i += 2) {
        double t1 = 1.0 - v(i);
        double t2 = 10.0 * (v(i + 1) - v(i) * v(i));
        g(i + 1) = 20.0 * t2;
        g(i) = -2.0 * (v(i) * g(i + 1) + t1);
        fx += t1 * t1 + t2 * t2;
    }
    return fx;
}
int iter_count;
int rosen_progress(void *, const Eigen::VectorXd &, const Eigen::VectorXd &, const double, const double, const int k, const int)
{
    iter_count = k;
    return 0;
}
} // namespace

extern "C" {

/* layouts as in oracle.h: inPs Dim x (N-1) col-major, head/tail Dim x 3 col-major, c / gdC 6N x Dim col-major */
int ref_minco(int Dim, int N, const double *inPs, const double *ts, const double *head, const double *tail, double *c, double *jerk,
              double *gdC_jerk, double *gdT_jerk, const double *gdC_in, double *gdT_io, double *gdP)
{
    if (Dim == 1) return minco_all<1>(N, inPs, ts, head, tail, c, jerk, gdC_jerk, gdT_jerk, gdC_in, gdT_io, gdP);
    if (Dim == 2) return minco_all<2>(N, inPs, ts, head, tail, c, jerk, gdC_jerk, gdT_jerk, gdC_in, gdT_io, gdP);
    return -1;
}

/* BandedSystem alone: A given densely (n x n, row-major, only the band is read), b n x m col-major; solve / adjoint solve in place */
int ref_banded(int n, int lo, int up, const double *A_dense, double *b, int m, int adjoint)
{
    BandedSystem A;
    A.create(n, lo, up);
    A.reset();
    for (int i = 0; i < n; i++)
        for (int j = std::max(0, i - lo); j <= std::min(n - 1, i + up); j++) A(i, j) = A_dense[(size_t)i * n + j];
    A.factorizeLU();
    Eigen::MatrixXd B(n, m);
    for (int cidx = 0; cidx < m; cidx++) for (int i = 0; i < n; i++) B(i, cidx) = b[i + (size_t)cidx * n];
    if (adjoint) A.solveAdj(B);
    else A.solve(B);
    for (int cidx = 0; cidx < m; cidx++) for (int i = 0; i < n; i++) b[i + (size_t)cidx * n] = B(i, cidx);
    A.destroy();
    return 0;
}

int ref_lbfgs_rosenbrock(int n, double *x, double *f, int mem_size, double g_epsilon, int past, double delta, int *iters, int *evals)
{
    Eigen::VectorXd xx(n);
    for (int i = 0; i < n; i++) xx(i) = x[i];
    lbfgs::lbfgs_parameter_t lp;
    lp.mem_size = mem_size; lp.g_epsilon = g_epsilon; lp.past = past; lp.delta = delta;
    Rosen inst{n, 0};
    double fx = 0.0;
    iter_count = 0;
    int r = lbfgs::lbfgs_optimize(xx, fx, rosen_eval, nullptr, rosen_progress, &inst, lp);
    for (int i = 0; i < n; i++) x[i] = xx(i);
    *f = fx;
    if (iters) *iters = iter_count;
    if (evals) *evals = inst.evals;
    return r;
}

} /* extern "C" */
