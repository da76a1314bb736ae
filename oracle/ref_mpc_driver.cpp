/* oracle/ref_mpc_driver.cpp -- the MPC side's own MINCO (mpc_controller/include/utils/minco_traj.hpp:336-460, compiled unmodified
 * from /root/reference against oracle/shim): TrajAnalyzer::setTraj (traj_anal.hpp:125-181) rebuilds the trajectory it tracks from
 * the SE2Traj message with MincoTraj<D>::reset(headState, N) + generate(inPs, tailState, ts).  tests/test_ref_pin.py checks that
 * the oracle's MINCO (and so the CUDA banded kernel) reproduces that re-solve bit for bit.  TEST INFRASTRUCTURE ONLY. */
#include "utils/minco_traj.hpp"

namespace {
template <int D>
int run(int N, const double *inPs, const double *ts, const double *head, const double *tail, double *c)
{
    mpc_utils::MincoTraj<D> mj;
    Eigen::MatrixXd H(D, 3), T(D, 3), P(D, N - 1);
    Eigen::VectorXd t(N);
    for (int j = 0; j < 3; j++) for (int d = 0; d < D; d++) { H(d, j) = head[d + j * D]; T(d, j) = tail[d + j * D]; }
    for (int j = 0; j < N - 1; j++) for (int d = 0; d < D; d++) P(d, j) = inPs[d + (size_t)j * D];
    for (int i = 0; i < N; i++) t(i) = ts[i];
    mj.reset(H, N);
    mj.generate(P, T, t);
    for (int d = 0; d < D; d++) for (int i = 0; i < 6 * N; i++) c[i + (size_t)d * 6 * N] = mj.b(i, d);
    return 0;
}
} // namespace

extern "C" int ref_mpc_minco(int Dim, int N, const double *inPs, const double *ts, const double *head, const double *tail, double *c)
{
    if (Dim == 1) return run<1>(N, inPs, ts, head, tail, c);
    if (Dim == 2) return run<2>(N, inPs, ts, head, tail, c);
    return -1;
}
