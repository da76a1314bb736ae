// ualm_tp_host.h -- internal C++ interface between the C ABI (ualm_api.cu) and the throughput engine (ualm_tp.cu).
// The engine is its own translation unit because it is compiled with FMA contraction on (the parity path needs -fmad=false).
#pragma once

#include <string>

#include "ualm.h"

namespace ualm_tp {

struct TpEngine;

// every function returns UALM_OK or a negative UALM_E* code; *err receives the message
int tp_create(TpEngine **e, int device, int precision, std::string *err);
void tp_destroy(TpEngine *e);
int tp_set_params(TpEngine *e, const ualm_params_t *p, std::string *err);
int tp_set_map(TpEngine *e, const ualm_map_geom_t *g, const float *cells, std::string *err);
// lanes = tickets: a lane holds one uploaded batch; admit puts it into the pool of running trajectories, collect drives the
// evaluation rounds until that lane's batch is finished and gathers its results on the device
int tp_upload(TpEngine *e, int lane, int B, const int32_t *N, const int32_t *M, const double *bnd, const double *total_time, const double *inner_xy,
              const double *inner_yaw, std::string *err);
int tp_admit(TpEngine *e, int lane, std::string *err);                       // asynchronous
int tp_collect(TpEngine *e, int lane, std::string *err);                     // blocks: rounds until the lane's batch is done
int tp_download(TpEngine *e, int lane, ualm_result_t *results, double *c_xy, double *c_yaw, std::string *err);
int tp_pack_records(TpEngine *e, int lane, double *d_records, int stride, std::string *err);
bool tp_lane_in_flight(TpEngine *e, int lane);
bool tp_lane_has_batch(TpEngine *e, int lane);
bool tp_lane_collected(TpEngine *e, int lane);
int tp_last_solve(TpEngine *e, int lane, float *ms, int *launches);
int tp_mark_begin(TpEngine *e, std::string *err);
int tp_mark_end(TpEngine *e, float *ms, std::string *err);
// kernel-level entry points over the lane's uploaded batch (mode 1: one innerCallback at the given x / duals; mode 2: initScaling)
int tp_eval(TpEngine *e, int lane, const double *x, const double *lambda, const double *mu, const double *scale_cx, const double *scale_fx, double rho,
            double *f, double *grad, double *hx, double *gx, double *c_xy, double *c_yaw, std::string *err);
int tp_init_scaling(TpEngine *e, int lane, double *scale_fx, double *scale_cx, std::string *err);
// time `reps` launches of kb_kernel alone over the lane's batch (the trajectories sit at their initial guess)
int tp_time_penalty(TpEngine *e, int lane, int reps, int use_tma, float *ms_per_launch, double *algorithmic_bytes, std::string *err);
// device pointers of the lane's gathered outputs (for the post-solve scan)
int tp_lane_outputs(TpEngine *e, int lane, const ualm_result_t **d_res, const double **d_cxy, const double **d_cyaw, int *B, const int32_t **N,
                    const int32_t **M);

} // namespace ualm_tp
