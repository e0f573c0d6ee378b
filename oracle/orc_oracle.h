/*
 * orc_oracle.h -- CPU ORACLE for the d2vins sliding-window BA hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load liborc_oracle.so.  Nothing under
 * d2slam_b200/ links, imports or calls it; the product path fails loudly without CUDA.
 *
 * PARITY: the factor level is PINNED to the reference's own classes -- oracle/_ref/libd2ref.so holds the unmodified
 * D2SLAM factor sources (projection*Factor.cpp, imu_factor.h + integration_base.h, consenus_factor.cpp,
 * pose_local_parameterization.cpp) compiled by oracle/Makefile.ref against the stand-in third-party headers of
 * oracle/_shim; tests/test_ref_pin.py compares every orc_*_eval with them and tests/golden/ref_factors.npz freezes their
 * outputs.  UNPINNED (restated from published algorithms, every such assumption marked ASSUMED in the source): the
 * arithmetic of ceres::Solve (ceres-solver, HKUST-Swarm fork, branch D2SLAM, nominal 2.1.0 -- docker/Dockerfile.x86:3,67),
 * swarm_msgs (Swarm::Pose), and the two pieces whose translation units cannot be compiled against stubs (loss corrector
 * BaseParamResInfo.cpp:71-92, PriorFactor prior_factor.cpp:45-177).  Those are checked against finite differences, the
 * optimality conditions of the same cost, and ConsensusSolver.cpp line by line.
 *
 * The oracle mirrors the C ABI of include/d2ba.h (same input structs, one window per
 * oracle handle) so parity tests drive both sides with identical calls.
 */
#ifndef ORC_ORACLE_H_
#define ORC_ORACLE_H_
#include <stdint.h>
#include "../include/d2ba.h"

#ifdef __cplusplus
extern "C" {
#endif

enum { ORC_PROJ_2F1C = 0, ORC_PROJ_2F2C = 1, ORC_PROJ_1F2C = 2, ORC_PROJ_2F1C_DEPTH = 3, ORC_PROJ_DEPTH_PRIOR = 4 };

typedef struct orc_obs_const {
  double pts_i[3], pts_j[3], vel_i[3], vel_j[3];
  double td_i, td_j;
  double tangent_base[6];
  double inv_depth_j; /* 1/depth of observation j (DEPTH type) */
} orc_obs_const;

typedef struct orc_imu_const {
  double sum_dt;
  double delta_p[3], delta_q[4], delta_v[3];
  double linearized_ba[3], linearized_bg[3];
  double jacobian[225], covariance[225], sqrt_info[225];
} orc_imu_const;

/* ---- factor level (orc_factors.c) ---- */
void orc_tangent_base(const double *pts_j, double *tb);
void orc_proj_eval(int type, const orc_obs_const *c, double sqrt_info_px, double depth_sqrt_inf,
                   const double *pose_i, const double *pose_j, const double *ext_a, const double *ext_b,
                   double inv_dep_i, double td, double *r, double *J_pose_i, double *J_pose_j,
                   double *J_ext_a, double *J_ext_b, double *J_lam, double *J_td);
void orc_depth_prior_eval(double inv_dep, double depth, double depth_sqrt_inf, double *r, double *J);
int orc_imu_sqrt_info(const double *cov, double *sqrt_info);
void orc_imu_residual(const orc_imu_const *p, double g_norm, const double *pose_i, const double *sb_i,
                      const double *pose_j, const double *sb_j, double *res15);
void orc_imu_eval(const orc_imu_const *p, double g_norm, const double *pose_i, const double *sb_i,
                  const double *pose_j, const double *sb_j, double *r, double *J_pose_i, double *J_sb_i,
                  double *J_pose_j, double *J_sb_j);
void orc_preintegrate(int n, const double *dt, const double *acc, const double *gyr, const double *ba,
                      const double *bg, double acc_n, double gyr_n, double acc_w, double gyr_w,
                      orc_imu_const *out);
void orc_consensus_eval(const double *t_ref, const double *q_ref_xyzw, const double *t_tilde,
                        const double *theta_tilde, double rho_T, double rho_theta, const double *pose,
                        double *r6, double *J6x7);
void orc_huber(double a, double s, double rho[3]);
void orc_corrector(const double rho[3], double sq_norm, double *residual_scaling, double *sqrt_rho1,
                   double *alpha_sq_norm);
void orc_pose_plus(const double *x, const double *delta, double *out);
void orc_prior_dx_pose(const double *x, const double *x0, double *dx6);
void orc_sym_eig(int n, const double *A, double *evals, double *V);
void orc_to_jac_res(int m, const double *A, const double *b, double *J, double *e0);
void orc_average_quats(int n, const double *q_xyzw, double *out_xyzw);
void orc_delta_pose_tangent(const double *a, const double *b, double *out6);

/* ---- solver level (orc_solver.c): mirrors include/d2ba.h, one window per handle ---- */
typedef struct orc_handle orc_handle;
int orc_create(const d2ba_config *cfg, orc_handle **out);
int orc_destroy(orc_handle *o);
int orc_reset(orc_handle *o);
int orc_set_blocks(orc_handle *o, int32_t kind, int32_t n, const int64_t *ids, const double *values,
                   const uint8_t *is_const);
int orc_add_proj(orc_handle *o, int32_t n, const d2ba_proj_obs *obs);
int orc_add_landmark_tracks(orc_handle *o, int32_t n_landmarks, const int64_t *landmark_ids,
                            const int32_t *track_ptr, const d2ba_track_obs *obs, int32_t fuse_dep,
                            double min_depth_to_fuse, double max_depth_to_fuse, int32_t n_ignore,
                            const int64_t *ignore_frames);
int orc_add_imu(orc_handle *o, int32_t n, const d2ba_imu *imu);
int orc_set_prior(orc_handle *o, int32_t m, const double *J, const double *e0, int32_t nblk,
                  const d2ba_blockref *refs, const double *x0);
int orc_set_prior_info(orc_handle *o, int32_t m, const double *A, const double *b, int32_t nblk,
                       const d2ba_blockref *refs, const double *x0);
int orc_set_consensus(orc_handle *o, int32_t n, const d2ba_blockref *refs, const int32_t *slot,
                      int32_t n_slots_global);
int orc_solve(orc_handle *o, d2ba_report *report);
int orc_solve_fixed(orc_handle *o, int32_t iters, d2ba_report *report);
/* ADMM over n agents in one process (in-memory all-gather, mirrors how
 * d2pgo/launch/d2pgo_test_multi.launch fakes the network). fixed_iters>0: no convergence exits. */
int orc_admm_solve(orc_handle **agents, int32_t n, int32_t fixed_mode, d2ba_report *reports);
/* n independent windows on nthreads host threads (1 thread per solve, like ceres num_threads=1) */
int orc_solve_many(orc_handle **hs, int32_t n, int32_t nthreads, int32_t fixed_iters, d2ba_report *reports);
int orc_admm_many(orc_handle **hs, int32_t n_swarms, int32_t n_agents, int32_t nthreads, int32_t fixed_mode, d2ba_report *reports);
int orc_get_blocks(orc_handle *o, int32_t kind, int32_t n, const int64_t *ids, double *out);
int orc_debug_linearize(orc_handle *o);
int orc_debug_get(orc_handle *o, int32_t item, void *out, int64_t out_bytes, int64_t *needed);
int orc_marginalize_x0(orc_handle *o, int32_t n_remove, const int64_t *remove_frame_ids, int32_t *m_out,
                       int32_t max_m, double *A_out, double *b_out, int32_t *nblk_out, int32_t max_blk,
                       d2ba_blockref *refs_out, double *x0_out);
/* consensus state access for tests */
int orc_get_consensus(orc_handle *o, int32_t n, const d2ba_blockref *refs, double *z7_out, double *tilde6_out);

#ifdef __cplusplus
}
#endif
#endif
