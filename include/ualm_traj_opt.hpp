// ualm_traj_opt.hpp -- C++ host-side mirror of the reference's back-end interface, over the C ABI of ualm.h.
//
// The reference class is uneven_planner::ALMTrajOpt (back_end/include/back_end/alm_traj_opt.h:21-120):
//     init(nh), setEnvironment(map), optimizeSE2Traj(initXY, endXY, innerXY, initYaw, endYaw, innerYaw, totalTime) -> int,
//     getTraj() -> SE2Trajectory
// This header keeps those names, argument meaning (column-major Eigen layouts) and return codes (0 ok / 1 L-BFGS error /
// 2 ALM max-iter, alm_traj_opt.cpp:176,252,267) so PlanManager::rcvWpsCallBack (plan_manager.cpp:134-138) compiles against
// it unchanged when Eigen is present (define UALM_WITH_EIGEN), and adds the batch entry point the GPU is built for.
// No ROS, no Eigen required: the plain-pointer overloads are always available.  Header-only; link libualm.so.
#pragma once

#include <stdexcept>
#include <string>
#include <vector>

#include "ualm.h"

#ifdef UALM_WITH_EIGEN
#include <Eigen/Eigen>
#endif

namespace uneven_planner_b200 {

// Result container with the reference's Piece / PolyTrajectory / SE2Trajectory conventions (se2traj.hpp:31-150, 255-414):
// per piece a duration and a Dim x 6 coefficient matrix, HIGHEST power first (MinJerkOpt::getTraj reverses the solver's
// low->high order, se2traj.hpp:682-695).
template <int Dim>
struct Piece {
    double duration = 0.0;
    double coeff[Dim][6];                       // coeff[d][0] * t^5 + ... + coeff[d][5]
    void getValue(double t, double *out) const  // Piece::getValue, se2traj.hpp:106-118
    {
        for (int d = 0; d < Dim; d++) {
            double v = 0.0, tn = 1.0;
            for (int i = 5; i >= 0; i--) { v += tn * coeff[d][i]; tn *= t; }
            out[d] = v;
        }
    }
    double getDuration() const { return duration; }
};

struct SE2Trajectory {
    std::vector<Piece<2>> pos_traj;
    std::vector<Piece<1>> yaw_traj;
    double getTotalDuration() const
    {
        double a = 0, b = 0;
        for (auto &p : pos_traj) a += p.duration;
        for (auto &p : yaw_traj) b += p.duration;
        return a < b ? a : b;                   // SE2Trajectory::getTotalDuration, se2traj.hpp:416-419
    }
};

// PolyTrajectory::locatePieceIdx + getValue (se2traj.hpp:343-367): piece containing time t (t is reduced to the piece-local time)
template <int Dim>
inline int locatePieceIdx(const std::vector<Piece<Dim>> &pieces, double &t)
{
    const int N = (int)pieces.size();
    int idx;
    double dur;
    for (idx = 0; idx < N && t > (dur = pieces[idx].getDuration()); idx++) t -= dur;
    if (idx == N) {
        idx--;
        t += pieces[idx].getDuration();
    }
    return idx;
}
template <int Dim>
inline void getValue(const std::vector<Piece<Dim>> &pieces, double t, double *out)
{
    const int idx = locatePieceIdx(pieces, t);
    pieces[idx].getValue(t, out);
}
template <int Dim>
inline double getTotalDuration(const std::vector<Piece<Dim>> &pieces)   // PolyTrajectory::getTotalDuration, se2traj.hpp:291-300
{
    double total = 0.0;
    for (const auto &p : pieces) total += p.getDuration();
    return total;
}

// What PlanManager publishes to the MPC (mpc_controller/msg/SE2Traj.msg; plan_manager.cpp:151-185): the piece start points,
// the end point and the piece durations of both splines; the boundary velocity / acceleration fields are zero in the reference.
struct SE2TrajMsg {
    std::vector<double> pos_pts;      // (N + 1) x 2, row-major: pos_traj[i].getValue(0), then pos_traj.getValue(total)
    std::vector<double> posT_pts;     // N
    std::vector<double> angle_pts;    // M + 1
    std::vector<double> angleT_pts;   // M
    double init_v[3] = {0, 0, 0}, init_a[3] = {0, 0, 0};
};
inline SE2TrajMsg toSE2TrajMsg(const SE2Trajectory &tr)
{
    SE2TrajMsg m;
    double p[2], a[1];
    for (const auto &pc : tr.pos_traj) {
        pc.getValue(0.0, p);
        m.pos_pts.push_back(p[0]); m.pos_pts.push_back(p[1]);
        m.posT_pts.push_back(pc.getDuration());
    }
    getValue(tr.pos_traj, getTotalDuration(tr.pos_traj), p);
    m.pos_pts.push_back(p[0]); m.pos_pts.push_back(p[1]);
    for (const auto &pc : tr.yaw_traj) {
        pc.getValue(0.0, a);
        m.angle_pts.push_back(a[0]);
        m.angleT_pts.push_back(pc.getDuration());
    }
    getValue(tr.yaw_traj, getTotalDuration(tr.yaw_traj), a);
    m.angle_pts.push_back(a[0]);
    return m;
}

// c_xy: 6N x 2 column-major, c_yaw: 6M (solver order, low -> high power); T_xy / T_yaw: the uniform piece durations T1(i) the
// reference's getTraj() carries (MinJerkOpt::getTraj, se2traj.hpp:682-695; calTfromTau, alm_traj_opt.h:257-261) --
// ualm_result_t::piece_T_xy / piece_T_yaw, NOT total_T / N (the N-fold sum divided back differs by a few ulps)
inline SE2Trajectory make_traj(int N, int M, const double *c_xy, const double *c_yaw, double T_xy, double T_yaw)
{
    SE2Trajectory tr;
    tr.pos_traj.resize(N);
    tr.yaw_traj.resize(M);
    for (int i = 0; i < N; i++) {
        tr.pos_traj[i].duration = T_xy;
        for (int d = 0; d < 2; d++)
            for (int k = 0; k < 6; k++) tr.pos_traj[i].coeff[d][5 - k] = c_xy[6 * i + k + d * 6 * N];
    }
    for (int i = 0; i < M; i++) {
        tr.yaw_traj[i].duration = T_yaw;
        for (int k = 0; k < 6; k++) tr.yaw_traj[i].coeff[0][5 - k] = c_yaw[6 * i + k];
    }
    return tr;
}
inline SE2Trajectory make_traj(int N, int M, const double *c_xy, const double *c_yaw, const ualm_result_t &r)
{
    return make_traj(N, M, c_xy, c_yaw, r.piece_T_xy, r.piece_T_yaw);
}

class ALMTrajOpt {
public:
    // the reference's public parameter members (alm_traj_opt.h:29-53) live in `params`
    ualm_params_t params;

    explicit ALMTrajOpt(int device = 0, int precision = 64)
    {
        ualm_default_params(&params);
        if (ualm_create(&ctx_, device, precision) != UALM_OK) throw std::runtime_error(std::string("ualm_create: ") + ualm_last_error());
    }
    ~ALMTrajOpt() { ualm_destroy(ctx_); }
    ALMTrajOpt(const ALMTrajOpt &) = delete;
    ALMTrajOpt &operator=(const ALMTrajOpt &) = delete;

    // ALMTrajOpt::init(nh): the caller fills `params` from its rosparam server, then calls init()
    void init() { check(ualm_set_params(ctx_, &params), "ualm_set_params"); }

    // ALMTrajOpt::setEnvironment(UnevenMap::Ptr): the map grid (UnevenMap::map_buffer is private in the reference,
    // uneven_map.h:91, so the maintainer-side binding passes geometry + a float4 view of the cells)
    void setEnvironment(const ualm_map_geom_t &geom, const float *cells_xyzw) { check(ualm_set_map(ctx_, &geom, cells_xyzw), "ualm_set_map"); }
    // the same from the reference's own double grid (UnevenMap::map_buffer, RXS2 = 4 doubles per cell, uneven_map.h:36-64)
    void setEnvironment(const ualm_map_geom_t &geom, const double *rxs2_cells, bool repack_to_float = false)
    {
        check(ualm_set_map_f64(ctx_, &geom, rxs2_cells, repack_to_float ? 1 : 0), "ualm_set_map_f64");
    }

    // ---- single problem, the reference's call (alm_traj_opt.h:92-98), plain pointers ----
    // initStateXY/endStateXY: 2x3 column-major; innerPtsXY: 2 x (N-1) column-major; initYaw/endYaw: 3; innerPtsYaw: M-1
    int optimizeSE2Traj(const double *initStateXY, const double *endStateXY, const double *innerPtsXY, int n_inner_xy,
                        const double *initYaw, const double *endYaw, const double *innerPtsYaw, int n_inner_yaw, double totalTime)
    {
        const int32_t N = n_inner_xy + 1, M = n_inner_yaw + 1;
        double bnd[18];
        for (int k = 0; k < 6; k++) { bnd[k] = initStateXY[k]; bnd[6 + k] = endStateXY[k]; }
        for (int k = 0; k < 3; k++) { bnd[12 + k] = initYaw[k]; bnd[15 + k] = endYaw[k]; }
        last_N_ = N; last_M_ = M;
        c_xy_.assign(12 * (size_t)N, 0.0);
        c_yaw_.assign(6 * (size_t)M, 0.0);
        check(ualm_solve_batch(ctx_, 1, &N, &M, bnd, &totalTime, innerPtsXY, innerPtsYaw, &last_, c_xy_.data(), c_yaw_.data()), "ualm_solve_batch");
        return last_.ret_code;
    }
    SE2Trajectory getTraj() const { return make_traj(last_N_, last_M_, c_xy_.data(), c_yaw_.data(), last_); }
    const ualm_result_t &lastResult() const { return last_; }

#ifdef UALM_WITH_EIGEN
    // the reference's exact signature (alm_traj_opt.h:92-98)
    int optimizeSE2Traj(const Eigen::MatrixXd &initStateXY, const Eigen::MatrixXd &endStateXY, const Eigen::MatrixXd &innerPtsXY,
                        const Eigen::VectorXd &initYaw, const Eigen::VectorXd &endYaw, const Eigen::VectorXd &innerPtsYaw,
                        const double &totalTime)
    {
        return optimizeSE2Traj(initStateXY.data(), endStateXY.data(), innerPtsXY.data(), (int)innerPtsXY.cols(), initYaw.data(),
                               endYaw.data(), innerPtsYaw.data(), (int)innerPtsYaw.size(), totalTime);
    }
#endif

    // ---- batch: B independent optimizeSE2Traj problems, packed as in ualm.h ----
    void optimizeBatch(int B, const int32_t *N, const int32_t *M, const double *bnd, const double *total_time, const double *inner_xy,
                       const double *inner_yaw, ualm_result_t *results, double *c_xy, double *c_yaw)
    {
        check(ualm_solve_batch(ctx_, B, N, M, bnd, total_time, inner_xy, inner_yaw, results, c_xy, c_yaw), "ualm_solve_batch");
    }

    // pipelined form: submit returns once the batch is uploaded and queued (up to `depth` batches in flight), wait collects it
    int submitBatch(int B, const int32_t *N, const int32_t *M, const double *bnd, const double *total_time, const double *inner_xy,
                    const double *inner_yaw, int depth = 2)
    {
        int ticket = -1;
        check(ualm_submit_batch(ctx_, B, N, M, bnd, total_time, inner_xy, inner_yaw, depth, &ticket), "ualm_submit_batch");
        return ticket;
    }
    void waitBatch(int ticket, ualm_result_t *results, double *c_xy, double *c_yaw) { check(ualm_wait_batch(ctx_, ticket, results, c_xy, c_yaw), "ualm_wait_batch"); }

    // What the MPC receives and tracks, for the batch solved last on the selected lane (ualm_mpc_export_batch): the SE2Traj message
    // arrays of PlanManager (plan_manager.cpp:151-185), the MPC side's own MINCO over them (traj_anal.hpp:125-181) and the largest
    // planned-vs-tracked deviation.  Any pointer may be null.
    void exportToMpcBatch(double dt, double *pos_pts, double *posT_pts, double *angle_pts, double *angleT_pts, double *c_mpc_xy, double *c_mpc_yaw,
                          double *dev4, const double *init_v = nullptr, const double *init_a = nullptr)
    {
        check(ualm_mpc_export_batch(ctx_, dt, init_v, init_a, pos_pts, posT_pts, angle_pts, angleT_pts, c_mpc_xy, c_mpc_yaw, dev4), "ualm_mpc_export_batch");
    }

    // The reference's whole chain for B (start, goal) pairs: KinoAstar::plan + PlanManager's resampler on host threads
    // (ualm_front_end_batch), then the batched optimizer.  packed[b] = index of pair b in the outputs or -1 (no path / over the
    // optimizer's limits); results / c_xy / c_yaw hold the packed problems back to back like optimizeBatch.  Returns their number.
    int planAndOptimizeBatch(const ualm_astar_map_t &map, const ualm_astar_params_t &ap, const ualm_resample_params_t &rp, int B, const double *starts,
                             const double *goals, std::vector<int32_t> &packed, std::vector<int32_t> &N, std::vector<int32_t> &M, std::vector<ualm_result_t> &results,
                             std::vector<double> &c_xy, std::vector<double> &c_yaw, int nthreads = 0)
    {
        packed.assign(B, -1); N.assign(B, 0); M.assign(B, 0);
        std::vector<double> bnd(18 * (size_t)B), T(B), ixy(2 * 63 * (size_t)B + 2), iyaw(127 * (size_t)B + 1);
        const int k = ualm_front_end_batch(&map, &ap, &rp, B, starts, goals, nthreads, N.data(), M.data(), bnd.data(), T.data(), ixy.data(), (long long)ixy.size(),
                                           iyaw.data(), (long long)iyaw.size(), packed.data(), nullptr);
        if (k < 0) throw std::runtime_error("ualm_front_end_batch failed");
        N.resize(k); M.resize(k);
        size_t ncx = 0, ncy = 0;
        for (int i = 0; i < k; i++) { ncx += 12 * (size_t)N[i]; ncy += 6 * (size_t)M[i]; }
        results.assign(k, ualm_result_t{}); c_xy.assign(ncx, 0.0); c_yaw.assign(ncy, 0.0);
        if (k > 0) optimizeBatch(k, N.data(), M.data(), bnd.data(), T.data(), ixy.data(), iyaw.data(), results.data(), c_xy.data(), c_yaw.data());
        return k;
    }

    ualm_ctx_t *handle() { return ctx_; }

private:
    static void check(int rc, const char *what)
    {
        if (rc != UALM_OK) throw std::runtime_error(std::string(what) + ": " + ualm_last_error());
    }
    ualm_ctx_t *ctx_ = nullptr;
    ualm_result_t last_{};
    int last_N_ = 0, last_M_ = 0;
    std::vector<double> c_xy_, c_yaw_;
};

} // namespace uneven_planner_b200
