// Drives the whole chain of PlanManager::rcvWpsCallBack for a batch through the C++ class (include/ualm_traj_opt.hpp):
// planAndOptimizeBatch (KinoAstar::plan + resampler on host threads, then the batched optimizer) and exportToMpcBatch.
// Input file: X Y W B | float cells | B x {start[3], goal[3]} doubles.  Output: one line per packed problem.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "ualm_traj_opt.hpp"

int main(int argc, char **argv)
{
    if (argc < 3) return 2;
    const int precision = atoi(argv[2]);
    FILE *f = fopen(argv[1], "rb");
    if (!f) return 3;
    int32_t hdr[4];
    if (fread(hdr, 4, 4, f) != 4) return 4;   // X, Y, W, B
    const size_t ncell = (size_t)hdr[0] * hdr[1] * hdr[2];
    const int B = hdr[3];
    std::vector<float> cells(4 * ncell);
    std::vector<double> sg(6 * (size_t)B);
    if (fread(cells.data(), 4, cells.size(), f) != cells.size()) return 5;
    if (fread(sg.data(), 8, sg.size(), f) != sg.size()) return 6;
    fclose(f);
    std::vector<double> starts(3 * (size_t)B), goals(3 * (size_t)B);
    for (int b = 0; b < B; b++)
        for (int k = 0; k < 3; k++) { starts[3 * b + k] = sg[6 * b + k]; goals[3 * b + k] = sg[6 * b + 3 + k]; }
    ualm_map_geom_t g;
    ualm_map_geometry(10.0, 10.0, 0.05, 0.1, &g);
    std::vector<uint8_t> occ3(ncell), occ2((size_t)hdr[0] * hdr[1]);
    if (ualm_map_occupancy(cells.data(), &g, 0.8, 0.003, occ3.data(), occ2.data()) != UALM_OK) return 7;
    ualm_astar_params_t ap;
    ualm_astar_default_params(&ap);
    ualm_resample_params_t rp{0.3, 2.0, 0.5, 1.2, 0.05};
    ualm_astar_map_t view{&g, cells.data(), nullptr, occ3.data(), occ2.data()};
    try {
        uneven_planner_b200::ALMTrajOpt opt(0, precision);
        opt.init();
        opt.setEnvironment(g, cells.data());
        std::vector<int32_t> packed, N, M;
        std::vector<ualm_result_t> res;
        std::vector<double> cxy, cyaw;
        const int k = opt.planAndOptimizeBatch(view, ap, rp, B, starts.data(), goals.data(), packed, N, M, res, cxy, cyaw, 2);
        printf("packed %d", k);
        for (int b = 0; b < B; b++) printf(" %d", packed[b]);
        printf("\n");
        size_t sN = 0, sM = 0;
        for (int i = 0; i < k; i++) { sN += N[i]; sM += M[i]; }
        std::vector<double> pos(2 * (sN + k)), posT(sN), ang(sM + k), angT(sM), mxy(12 * sN), myaw(6 * sM), dev(4 * (size_t)k);
        opt.exportToMpcBatch(0.01, pos.data(), posT.data(), ang.data(), angT.data(), mxy.data(), myaw.data(), dev.data());
        size_t ocx = 0, on = 0;
        for (int i = 0; i < k; i++) {
            printf("%d %d %d %d %.17g %.17g %.17g %.17g %.17g\n", N[i], M[i], res[i].ret_code, res[i].n_evals, res[i].inner_cost, cxy[ocx + 1], pos[2 * (on + i)],
                   mxy[ocx + 3], dev[4 * i]);
            ocx += 12 * (size_t)N[i]; on += N[i];
        }
    } catch (const std::exception &e) {
        printf("EXCEPTION %s\n", e.what());
        return 10;
    }
    return 0;
}
