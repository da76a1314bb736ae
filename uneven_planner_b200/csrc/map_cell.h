// map_cell.h -- one UnevenMap cell (x, y, yaw) -> (z, sigma, zb.x, zb.y), shared by the host builder (host_tools.cpp,
// ualm_map_build) and the CUDA builder (ualm_api.cu, ualm_map_build_device) so that both run the SAME arithmetic and produce
// bit-identical grids (IEEE double/float, no contraction: nvcc -fmad=false, gcc -ffp-contract=off; sin/cos from ualm_detmath.h).
//
// Restates UnevenMap::constructMap's per-cell loop and UnevenMap::filter (uneven_map/src/uneven_map.cpp:317-398, 5-43) over a
// uniform XY bin grid of the preprocessed cloud instead of PCL kd-trees; shares no code with them.
#pragma once

#include <math.h>
#include <stdint.h>

#include "ualm.h"
#include "ualm_detmath.h"

struct UalmMapPrep {
    const float *pts;     // bin-sorted cloud, xyz interleaved
    const int *start;     // nx*ny+1 bin offsets (bin = ix * ny + iy)
    int npts, nx, ny;
    double x0, y0, inv;   // bin origin and 1 / bin size
    double box_r;         // max ellipsoid semi-axis (uneven_map.cpp:319) = bin size
    double einv[3];       // 1 / ellipsoid semi-axes
    int iter_num;
};

UALM_HD int ualm_bin_x(const UalmMapPrep &g, double x) { return (int)floor((x - g.x0) * g.inv); }
UALM_HD int ualm_bin_y(const UalmMapPrep &g, double y) { return (int)floor((y - g.y0) * g.inv); }

// symmetric 3x3 eigen-decomposition, cyclic Jacobi.  a is destroyed; w = eigenvalues, v columns = eigenvectors
UALM_HD void ualm_jacobi3(double a[3][3], double w[3], double v[3][3])
{
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) v[i][j] = (i == j) ? 1.0 : 0.0;
    for (int sweep = 0; sweep < 32; sweep++) {
        const double off = fabs(a[0][1]) + fabs(a[0][2]) + fabs(a[1][2]);
        const double diag = fabs(a[0][0]) + fabs(a[1][1]) + fabs(a[2][2]);
        if (off <= 1e-300 || off <= 1e-18 * diag) break;
        for (int p = 0; p < 2; p++)
            for (int q = p + 1; q < 3; q++) {
                if (a[p][q] == 0.0) continue;
                const double theta = (a[q][q] - a[p][p]) / (2.0 * a[p][q]);
                const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
                for (int k = 0; k < 3; k++) { // A <- A J
                    const double akp = a[k][p], akq = a[k][q];
                    a[k][p] = c * akp - s * akq;
                    a[k][q] = s * akp + c * akq;
                }
                for (int k = 0; k < 3; k++) { // A <- J^T A
                    const double apk = a[p][k], aqk = a[q][k];
                    a[p][k] = c * apk - s * aqk;
                    a[q][k] = s * apk + c * aqk;
                }
                for (int k = 0; k < 3; k++) {
                    const double vkp = v[k][p], vkq = v[k][q];
                    v[k][p] = c * vkp - s * vkq;
                    v[k][q] = s * vkp + c * vkq;
                }
            }
    }
    for (int i = 0; i < 3; i++) w[i] = a[i][i];
}

// cell (x, y, w) of the grid `geom` -> out[4] = {z, sigma, zb.x, zb.y}
UALM_HD void ualm_map_cell(const UalmMapPrep &g, const ualm_map_geom_t &geom, int x, int y, int w, float out[4])
{
    double z = 0.0, sigma = 0.0, zbx = 0.0, zby = 0.0, cc = 1.0; // RXS2(), c_buffer = 1 (uneven_map.cpp:118-119)
    const double px = (x + 0.5) * geom.xy_resolution + geom.origin[0]; // indexToPos
    const double py = (y + 0.5) * geom.xy_resolution + geom.origin[1];
    const double pyaw = (w + 0.5) * geom.yaw_resolution + geom.origin[2];
    double syaw, cyaw;
    ualm_sincos(pyaw, &syaw, &cyaw);
    for (int iter = 0; iter < g.iter_num; iter++) { // uneven_map.cpp:326-398
        const double xyaw[3] = {cyaw, syaw, 0.0};
        const double zb[3] = {zbx, zby, cc};
        double yb[3] = {zb[1] * xyaw[2] - zb[2] * xyaw[1], zb[2] * xyaw[0] - zb[0] * xyaw[2], zb[0] * xyaw[1] - zb[1] * xyaw[0]};
        const double nyb = sqrt(yb[0] * yb[0] + yb[1] * yb[1] + yb[2] * yb[2]);
        if (nyb > 0) { yb[0] /= nyb; yb[1] /= nyb; yb[2] /= nyb; }
        const double xb[3] = {yb[1] * zb[2] - yb[2] * zb[1], yb[2] * zb[0] - yb[0] * zb[2], yb[0] * zb[1] - yb[1] * zb[0]};
        double wp[3] = {px + xb[0] * 0.12, py + xb[1] * 0.12, z};
        if (iter == 0 && g.npts > 0) { // nearest cloud point in the XY plane (uneven_map.cpp:346-355)
            const float qx = (float)wp[0], qy = (float)wp[1];
            int bx = ualm_bin_x(g, qx), by = ualm_bin_y(g, qy);
            bx = bx < 0 ? 0 : (bx > g.nx - 1 ? g.nx - 1 : bx);
            by = by < 0 ? 0 : (by > g.ny - 1 ? g.ny - 1 : by);
            float best = 1e30f, bestz = 0;
            const int rmax = g.nx > g.ny ? g.nx : g.ny;
            for (int ring = 0; ring < rmax; ring++) {
                for (int ix = bx - ring; ix <= bx + ring; ix++) {
                    if (ix < 0 || ix >= g.nx) continue;
                    for (int iy = by - ring; iy <= by + ring; iy++) {
                        if (iy < 0 || iy >= g.ny) continue;
                        const int ax = ix - bx < 0 ? bx - ix : ix - bx, ay = iy - by < 0 ? by - iy : iy - by;
                        if ((ax > ay ? ax : ay) != ring) continue;
                        for (int q = g.start[ix * g.ny + iy]; q < g.start[ix * g.ny + iy + 1]; q++) {
                            const float *p = g.pts + 3 * (size_t)q;
                            const float d = (p[0] - qx) * (p[0] - qx) + (p[1] - qy) * (p[1] - qy);
                            if (d < best) { best = d; bestz = p[2]; }
                        }
                    }
                }
                // every unvisited point is at least ring*bin away (query clamped into the grid)
                const double reach = (double)ring * g.box_r;
                if (best < 1e29f && (double)best <= reach * reach) break;
            }
            if (best < 1e29f) wp[2] = bestz;
        }
        // points inside the robot-frame ellipsoid (uneven_map.cpp:357-378): pass 0 = count and mean, pass 1 = covariance.  The
        // membership test is recomputed in the second pass (same operands, same result) instead of keeping a list.
        int bx0 = ualm_bin_x(g, wp[0] - g.box_r), bx1 = ualm_bin_x(g, wp[0] + g.box_r);
        int by0 = ualm_bin_y(g, wp[1] - g.box_r), by1 = ualm_bin_y(g, wp[1] + g.box_r);
        if (bx0 < 0) bx0 = 0;
        if (by0 < 0) by0 = 0;
        if (bx1 > g.nx - 1) bx1 = g.nx - 1;
        if (by1 > g.ny - 1) by1 = g.ny - 1;
        double m[3] = {0, 0, 0}, n = 0.0;
        double cov[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
        for (int pass = 0; pass < 2; pass++) {
            if (pass == 1) {
                if (n == 0.0) break;
                m[0] /= n; m[1] /= n; m[2] /= n;
            }
            for (int ix = bx0; ix <= bx1; ix++)
                for (int iy = by0; iy <= by1; iy++)
                    for (int q = g.start[ix * g.ny + iy]; q < g.start[ix * g.ny + iy + 1]; q++) {
                        const float *p = g.pts + 3 * (size_t)q;
                        const double d[3] = {p[0] - wp[0], p[1] - wp[1], p[2] - wp[2]};
                        if (d[0] * d[0] + d[1] * d[1] + d[2] * d[2] > g.box_r * g.box_r * 1.0001) continue;
                        const double r0 = (xb[0] * d[0] + xb[1] * d[1] + xb[2] * d[2]) * g.einv[0];
                        const double r1 = (yb[0] * d[0] + yb[1] * d[1] + yb[2] * d[2]) * g.einv[1];
                        const double r2 = (zb[0] * d[0] + zb[1] * d[1] + zb[2] * d[2]) * g.einv[2];
                        if (!(r0 * r0 + r1 * r1 + r2 * r2 < 1.0)) continue;
                        if (pass == 0) {
                            m[0] += p[0]; m[1] += p[1]; m[2] += p[2];
                            n += 1.0;
                        } else {
                            const double v[3] = {p[0] - m[0], p[1] - m[1], p[2] - m[2]};
                            for (int a = 0; a < 3; a++) for (int b = 0; b < 3; b++) cov[a][b] += v[a] * v[b];
                        }
                    }
        }
        if (n == 0.0) { // uneven_map.cpp:379-386
            z = wp[2]; sigma = 0.0; zbx = 0.0; zby = 0.0; cc = 1.0;
        } else { // UnevenMap::filter (uneven_map.cpp:5-43)
            for (int a = 0; a < 3; a++) for (int b = 0; b < 3; b++) cov[a][b] /= n;
            double wv[3], ev[3][3];
            ualm_jacobi3(cov, wv, ev);
            int k = 0;
            if (wv[1] < wv[k]) k = 1;
            if (wv[2] < wv[k]) k = 2;
            double nn[3] = {ev[0][k], ev[1][k], ev[2][k]};
            const double nl = sqrt(nn[0] * nn[0] + nn[1] * nn[1] + nn[2] * nn[2]);
            nn[0] /= nl; nn[1] /= nl; nn[2] /= nl;
            if (nn[2] < 0.0) { nn[0] = -nn[0]; nn[1] = -nn[1]; nn[2] = -nn[2]; }
            double sg = wv[k] / (wv[0] + wv[1] + wv[2]) * 3.0;
            if (sg != sg) { sg = 1.0; nn[0] = 1.0; nn[1] = 0.0; nn[2] = 0.0; }
            z = m[2]; sigma = sg; zbx = nn[0]; zby = nn[1];
            cc = sqrt(1.0 - zbx * zbx - zby * zby);
        }
    }
    out[0] = (float)z; out[1] = (float)sigma; out[2] = (float)zbx; out[3] = (float)zby;
}
