// Drives the C++ ALMTrajOpt mirror (include/ualm_traj_opt.hpp) the way PlanManager drives the reference:
// init -> setEnvironment -> optimizeSE2Traj -> getTraj.  Reads one problem + map from a binary file written by the test.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "ualm_traj_opt.hpp"

int main(int argc, char **argv)
{
    if (argc < 2) return 2;
    FILE *f = fopen(argv[1], "rb");
    if (!f) return 3;
    int32_t hdr[5];
    if (fread(hdr, 4, 5, f) != 5) return 4;   // X, Y, W, N, M
    const size_t ncell = (size_t)hdr[0] * hdr[1] * hdr[2];
    std::vector<float> cells(4 * ncell);
    std::vector<double> bnd(18), ixy(2 * (hdr[3] - 1)), iyaw(hdr[4] - 1);
    double T;
    if (fread(cells.data(), 4, cells.size(), f) != cells.size()) return 5;
    if (fread(bnd.data(), 8, 18, f) != 18 || fread(&T, 8, 1, f) != 1) return 6;
    if (fread(ixy.data(), 8, ixy.size(), f) != ixy.size() || fread(iyaw.data(), 8, iyaw.size(), f) != iyaw.size()) return 7;
    fclose(f);
    try {
        uneven_planner_b200::ALMTrajOpt opt(0, 64);
        opt.init();
        ualm_map_geom_t g;
        ualm_map_geometry(10.0, 10.0, 0.05, 0.1, &g);
        opt.setEnvironment(g, cells.data());
        int ret = opt.optimizeSE2Traj(&bnd[0], &bnd[6], ixy.data(), hdr[3] - 1, &bnd[12], &bnd[15], iyaw.data(), hdr[4] - 1, T);
        auto tr = opt.getTraj();
        printf("ret %d evals %d cost %.17g T %.17g\n", ret, opt.lastResult().n_evals, opt.lastResult().inner_cost, tr.getTotalDuration());
        double p[2];
        tr.pos_traj[0].getValue(0.0, p);
        printf("start %.17g %.17g\n", p[0], p[1]);
        printf("durations %.17g %.17g\n", tr.pos_traj[0].getDuration(), tr.yaw_traj[0].getDuration());
        for (size_t i = 0; i < tr.pos_traj.size(); i++)
            for (int d = 0; d < 2; d++)
                for (int k = 0; k < 6; k++) printf("%.17g\n", tr.pos_traj[i].coeff[d][k]);
    } catch (const std::exception &e) {
        printf("EXCEPTION %s\n", e.what());
        return 10;
    }
    return 0;
}
