// Host-only check of the result containers of include/ualm_traj_opt.hpp (no CUDA device needed): coefficient order of
// make_traj (se2traj.hpp:682-695), Piece::getValue, locatePieceIdx and the SE2Traj export of plan_manager.cpp:151-185.
#include <cmath>
#include <cstdio>
#include <vector>
#include "ualm_traj_opt.hpp"

using namespace uneven_planner_b200;

static int fails = 0;
#define CHECK(cond) do { if (!(cond)) { printf("FAIL line %d: %s\n", __LINE__, #cond); fails++; } } while (0)

int main()
{
    // two xy pieces, three yaw pieces, total duration 3 s; x(t) = 1 + 2t (continuous over the pieces), y = t^2 on piece 0
    const int N = 2, M = 3;
    const double T = 3.0;
    std::vector<double> cxy(12 * N, 0.0), cyaw(6 * M, 0.0);
    // column-major 6N x 2: x column then y column; piece-local polynomials, low -> high power
    cxy[0] = 1.0; cxy[1] = 2.0;                    // piece 0: x = 1 + 2 s
    cxy[6] = 1.0 + 2.0 * 1.5; cxy[7] = 2.0;        // piece 1: x = 4 + 2 s
    cxy[12 + 2] = 1.0;                             // piece 0: y = s^2
    cxy[12 + 6] = 2.25; cxy[12 + 7] = 3.0; cxy[12 + 8] = 1.0;   // piece 1: y = (1.5 + s)^2
    for (int i = 0; i < M; i++) { cyaw[6 * i] = 0.5 * i; cyaw[6 * i + 1] = 0.5; }   // yaw = 0.5 t, piece duration 1
    SE2Trajectory tr = make_traj(N, M, cxy.data(), cyaw.data(), T / N, T / M);   // piece durations 1.5 and 1 (exact)
    CHECK(tr.pos_traj.size() == 2 && tr.yaw_traj.size() == 3);
    CHECK(tr.pos_traj[0].getDuration() == 1.5 && tr.yaw_traj[0].getDuration() == 1.0);
    CHECK(tr.pos_traj[0].coeff[0][5] == 1.0 && tr.pos_traj[0].coeff[0][4] == 2.0);      // highest power first
    CHECK(tr.getTotalDuration() == 3.0);
    double p[2], a[1];
    getValue(tr.pos_traj, 2.0, p);
    CHECK(std::fabs(p[0] - 5.0) < 1e-14 && std::fabs(p[1] - 4.0) < 1e-14);
    double t = 2.0;
    CHECK(locatePieceIdx(tr.pos_traj, t) == 1 && t == 0.5);
    t = 7.0;                                        // past the end: last piece, local time beyond its duration (se2traj.hpp:355-359)
    CHECK(locatePieceIdx(tr.pos_traj, t) == 1 && t == 5.5);
    getValue(tr.yaw_traj, 3.0, a);
    CHECK(std::fabs(a[0] - 1.5) < 1e-14);
    SE2TrajMsg m = toSE2TrajMsg(tr);
    CHECK(m.pos_pts.size() == 2 * (N + 1) && m.posT_pts.size() == N && m.angle_pts.size() == M + 1 && m.angleT_pts.size() == M);
    CHECK(m.pos_pts[0] == 1.0 && m.pos_pts[1] == 0.0 && m.pos_pts[2] == 4.0 && m.pos_pts[3] == 2.25);
    CHECK(std::fabs(m.pos_pts[4] - 7.0) < 1e-14 && std::fabs(m.pos_pts[5] - 9.0) < 1e-14);
    CHECK(m.angle_pts[0] == 0.0 && m.angle_pts[1] == 0.5 && m.angle_pts[2] == 1.0 && std::fabs(m.angle_pts[3] - 1.5) < 1e-14);
    CHECK(m.posT_pts[0] == 1.5 && m.angleT_pts[2] == 1.0 && m.init_v[0] == 0.0 && m.init_a[2] == 0.0);
    printf(fails ? "FAILED %d\n" : "OK\n", fails);
    return fails ? 1 : 0;
}
