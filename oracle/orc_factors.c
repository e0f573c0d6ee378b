/*
 * orc_factors.c -- CPU ORACLE (test infrastructure, never shipped, never timed as product).
 *
 * Plain-C restatement of every cost function on the d2vins bundle-adjustment hot path.
 * Parity status: UNPINNED at the Ceres boundary -- the reference ships no golden vectors
 * for this path and cannot be compiled here (no Eigen/Ceres/ROS);
 * each function is pinned instead by finite-difference checks (tests/test_oracle_factors.py),
 * the same scheme as the reference's own check() routines
 * (d2vins/src/factors/projectionTwoFrameOneCamFactor.cpp:179-305).
 *
 * Jacobian layout follows the reference exactly: row-major, ambient width (7 for pose /
 * extrinsic blocks with a zero 7th column), so these can be compared 1:1 with
 * ceres::CostFunction::Evaluate outputs when a Ceres build is available.
 */
#include "orc_oracle.h"
#include "orc_math.h"
#include <stdlib.h>

/* tangent base of the unit-sphere residual.
 * Follows d2vins/src/factors/projectionTwoFrameOneCamFactor.cpp:34-45. */
void orc_tangent_base(const double *pts_j, double *tb /*2x3*/) {
  double n = v3_norm(pts_j), a[3] = {pts_j[0] / n, pts_j[1] / n, pts_j[2] / n};
  double tmp[3] = {0, 0, 1};
  if (a[0] == tmp[0] && a[1] == tmp[1] && a[2] == tmp[2]) { tmp[0] = 1; tmp[1] = 0; tmp[2] = 0; }
  double d = v3_dot(a, tmp), b1[3] = {tmp[0] - a[0] * d, tmp[1] - a[1] * d, tmp[2] - a[2] * d};
  double n1 = v3_norm(b1);
  b1[0] /= n1; b1[1] /= n1; b1[2] /= n1;
  double b2[3];
  v3_cross(b2, a, b1);
  tb[0] = b1[0]; tb[1] = b1[1]; tb[2] = b1[2];
  tb[3] = b2[0]; tb[4] = b2[1]; tb[5] = b2[2];
}

/* d(normalize)/dx at x: I/|x| - x x^T/|x|^3 (norm_jaco, cpp:98-108).  The reference's
 * reduce_j_td uses the *norm of pts_camera_j* in the diagonal term but pts_j_td in the
 * outer product (cpp:111-118) -- reproduced by passing norm separately. */
static void norm_jaco(double *o, const double *x, double norm_diag) {
  double n3 = pow(v3_norm(x), 3);
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) o[i * 3 + j] = (i == j ? 1.0 / norm_diag : 0.0) - x[i] * x[j] / n3;
}

/* o(rows x 6) = reduce(rows x 3) * [L(3x3) | R(3x3)] written into a rows x 7 row-major block */
static void put_block7(double *J, int rows, const double *reduce, const double *L, const double *R) {
  for (int r = 0; r < rows; r++) {
    for (int c = 0; c < 3; c++) {
      double sl = 0, sr = 0;
      for (int k = 0; k < 3; k++) { sl += reduce[r * 3 + k] * L[k * 3 + c]; sr += reduce[r * 3 + k] * R[k * 3 + c]; }
      J[r * 7 + c] = sl;
      J[r * 7 + 3 + c] = sr;
    }
    J[r * 7 + 6] = 0.0;
  }
}

/*
 * Unified evaluation of the four two-view reprojection factors.
 *   type 2F1C : d2vins/src/factors/projectionTwoFrameOneCamFactor.cpp:48-177
 *   type 2F2C : d2vins/src/factors/projectionTwoFrameTwoCamFactor.cpp:46-187
 *   type 1F2C : d2vins/src/factors/projectionOneFrameTwoCamFactor.cpp:46-160
 *   type 2F1C_DEPTH : d2vins/src/factors/projectionTwoFrameOneCamDepthFactor.cpp:47-186
 * UNIT_SPHERE_ERROR is defined in the reference build (d2vins/src/d2vins_params.hpp:8).
 * Inputs are the parameter blocks in the reference's order; unused ones may be NULL.
 * Outputs: r (2 or 3), Jacobians row-major (rows x 7 / rows x 1); any may be NULL.
 */
void orc_proj_eval(int type, const orc_obs_const *c, double sqrt_info_px, double depth_sqrt_inf,
                   const double *pose_i, const double *pose_j, const double *ext_a, const double *ext_b,
                   double inv_dep_i, double td, double *r, double *J_pose_i, double *J_pose_j,
                   double *J_ext_a, double *J_ext_b, double *J_lam, double *J_td) {
  const int rows = (type == ORC_PROJ_2F1C_DEPTH) ? 3 : 2;
  double Pi[3] = {0, 0, 0}, Pj[3] = {0, 0, 0};
  oq_t Qi = {0, 0, 0, 1}, Qj = {0, 0, 0, 1};
  if (type != ORC_PROJ_1F2C) {
    v3_cpy(Pi, pose_i); Qi = q_from(pose_i + 3);
    v3_cpy(Pj, pose_j); Qj = q_from(pose_j + 3);
  }
  double tic[3], tic2[3];
  v3_cpy(tic, ext_a);
  oq_t qic = q_from(ext_a + 3), qic2;
  if (type == ORC_PROJ_2F2C || type == ORC_PROJ_1F2C) { v3_cpy(tic2, ext_b); qic2 = q_from(ext_b + 3); }
  else { v3_cpy(tic2, tic); qic2 = qic; }

  double pts_i_td[3], pts_j_td[3];
  for (int k = 0; k < 3; k++) {
    pts_i_td[k] = c->pts_i[k] - (td - c->td_i) * c->vel_i[k];
    pts_j_td[k] = c->pts_j[k] - (td - c->td_j) * c->vel_j[k];
  }
  double pts_camera_i[3] = {pts_i_td[0] / inv_dep_i, pts_i_td[1] / inv_dep_i, pts_i_td[2] / inv_dep_i};
  double pts_imu_i[3], pts_w[3], pts_imu_j[3], pts_camera_j[3], tmp[3];
  q_rot(pts_imu_i, qic, pts_camera_i); v3_add(pts_imu_i, pts_imu_i, tic);
  if (type == ORC_PROJ_1F2C) {
    v3_cpy(pts_imu_j, pts_imu_i);
  } else {
    q_rot(pts_w, Qi, pts_imu_i); v3_add(pts_w, pts_w, Pi);
    v3_sub(tmp, pts_w, Pj); q_rot(pts_imu_j, q_inv(Qj), tmp);
  }
  v3_sub(tmp, pts_imu_j, tic2); q_rot(pts_camera_j, q_inv(qic2), tmp);

  double ncj = v3_norm(pts_camera_j), njt = v3_norm(pts_j_td);
  double e[3] = {pts_camera_j[0] / ncj - pts_j_td[0] / njt, pts_camera_j[1] / ncj - pts_j_td[1] / njt,
                 pts_camera_j[2] / ncj - pts_j_td[2] / njt};
  const double *tb = c->tangent_base;
  double res[3];
  res[0] = tb[0] * e[0] + tb[1] * e[1] + tb[2] * e[2];
  res[1] = tb[3] * e[0] + tb[4] * e[1] + tb[5] * e[2];
  double sq[3] = {sqrt_info_px, sqrt_info_px, depth_sqrt_inf};
  if (rows == 3) res[2] = 1.0 / ncj - c->inv_depth_j;
  for (int k = 0; k < rows; k++) r[k] = sq[k] * res[k];

  if (!J_pose_i && !J_pose_j && !J_ext_a && !J_ext_b && !J_lam && !J_td) return;

  double Ri[9], Rj[9], ric[9], ric2[9], ric2_t[9], Rj_t[9];
  q_to_R(Ri, Qi); q_to_R(Rj, Qj); q_to_R(ric, qic); q_to_R(ric2, qic2);
  m3_T(ric2_t, ric2); m3_T(Rj_t, Rj);
  double J_w[9], J_imu_i[9], J_cam_i[9];
  if (type == ORC_PROJ_1F2C) {
    /* ric2_t_ric plays the role of J_cam_i (cpp:90) */
    m3_mul(J_cam_i, ric2_t, ric);
    memset(J_w, 0, sizeof J_w); memset(J_imu_i, 0, sizeof J_imu_i);
  } else {
    m3_mul(J_w, ric2_t, Rj_t);
    m3_mul(J_imu_i, J_w, Ri);
    m3_mul(J_cam_i, J_imu_i, ric);
  }
  double nj[9], reduce[9] = {0}, reduce_j_td[9];
  norm_jaco(nj, pts_camera_j, ncj);
  mm(reduce, tb, nj, 2, 3, 3);
  if (rows == 3) {
    double n3 = pow(ncj, 3);
    reduce[6] = -pts_camera_j[0] / n3; reduce[7] = -pts_camera_j[1] / n3; reduce[8] = -pts_camera_j[2] / n3;
  }
  norm_jaco(reduce_j_td, pts_j_td, ncj); /* sic: diagonal uses |pts_camera_j| */
  for (int rr = 0; rr < rows; rr++)
    for (int k = 0; k < 3; k++) reduce[rr * 3 + k] *= sq[rr];

  double S[9], L[9], R[9], neg[9];
  if (type != ORC_PROJ_1F2C) {
    if (J_pose_i) {
      m3_skew(S, pts_imu_i); m3_scale(neg, S, -1.0); m3_mul(R, J_imu_i, neg);
      put_block7(J_pose_i, rows, reduce, J_w, R);
    }
    if (J_pose_j) {
      m3_scale(L, J_w, -1.0); m3_skew(S, pts_imu_j); m3_mul(R, ric2_t, S);
      put_block7(J_pose_j, rows, reduce, L, R);
    }
  }
  if (type == ORC_PROJ_2F1C || type == ORC_PROJ_2F1C_DEPTH) {
    if (J_ext_a) {
      double ric_t[9]; m3_T(ric_t, ric);
      m3_subm(L, J_imu_i, ric_t);
      double t1[9], t2[9], t3[9], v[3], w1[3], w2[3];
      m3_skew(S, pts_camera_i); m3_mul(t1, J_cam_i, S); m3_scale(t1, t1, -1.0);
      m3_vec(v, J_cam_i, pts_camera_i); m3_skew(t2, v);
      m3_vec(w1, Ri, tic); v3_add(w1, w1, Pi); v3_sub(w1, w1, Pj); m3_vec(w1, J_w, w1);
      m3_vec(w2, Rj_t, tic); v3_sub(w1, w1, w2); m3_skew(t3, w1);
      m3_addm(R, t1, t2); m3_addm(R, R, t3);
      put_block7(J_ext_a, rows, reduce, L, R);
    }
  } else if (type == ORC_PROJ_2F2C) {
    if (J_ext_a) {
      m3_skew(S, pts_camera_i); m3_scale(neg, S, -1.0); m3_mul(R, J_cam_i, neg);
      put_block7(J_ext_a, rows, reduce, J_imu_i, R);
    }
    if (J_ext_b) {
      m3_scale(L, ric2_t, -1.0); m3_skew(R, pts_camera_j);
      put_block7(J_ext_b, rows, reduce, L, R);
    }
  } else { /* 1F2C */
    if (J_ext_a) {
      m3_skew(S, pts_camera_i); m3_scale(neg, S, -1.0); m3_mul(R, J_cam_i, neg);
      put_block7(J_ext_a, rows, reduce, ric2_t, R);
    }
    if (J_ext_b) {
      m3_scale(L, ric2_t, -1.0); m3_skew(R, pts_camera_j);
      put_block7(J_ext_b, rows, reduce, L, R);
    }
  }
  if (J_lam) {
    double v[3];
    m3_vec(v, J_cam_i, pts_i_td);
    for (int rr = 0; rr < rows; rr++)
      J_lam[rr] = (reduce[rr * 3] * v[0] + reduce[rr * 3 + 1] * v[1] + reduce[rr * 3 + 2] * v[2]) * -1.0 / (inv_dep_i * inv_dep_i);
  }
  if (J_td) {
    double v[3], w[3], t2[2];
    m3_vec(v, J_cam_i, c->vel_i);
    m3_vec(w, reduce_j_td, c->vel_j);
    t2[0] = sqrt_info_px * (tb[0] * w[0] + tb[1] * w[1] + tb[2] * w[2]);
    t2[1] = sqrt_info_px * (tb[3] * w[0] + tb[4] * w[1] + tb[5] * w[2]);
    for (int rr = 0; rr < rows; rr++) {
      double a = (reduce[rr * 3] * v[0] + reduce[rr * 3 + 1] * v[1] + reduce[rr * 3 + 2] * v[2]) / inv_dep_i * -1.0;
      J_td[rr] = a + (rr < 2 ? t2[rr] : 0.0);
    }
  }
}

/* OneFrameDepth, d2vins/src/factors/depth_factor.h:9-29 (autodiff of a linear function) */
void orc_depth_prior_eval(double inv_dep, double depth, double depth_sqrt_inf, double *r, double *J) {
  r[0] = (inv_dep - 1.0 / depth) * depth_sqrt_inf;
  if (J) J[0] = depth_sqrt_inf;
}

/* ------------------------------------------------------------------ IMU */
static void chol_lower(double *A, int n, int *ok) { /* in place, row-major, lower */
  *ok = 1;
  for (int j = 0; j < n; j++) {
    double d = A[j * n + j];
    for (int k = 0; k < j; k++) d -= A[j * n + k] * A[j * n + k];
    if (!(d > 0)) { *ok = 0; return; }
    d = sqrt(d);
    A[j * n + j] = d;
    for (int i = j + 1; i < n; i++) {
      double s = A[i * n + j];
      for (int k = 0; k < j; k++) s -= A[i * n + k] * A[j * n + k];
      A[i * n + j] = s / d;
    }
    for (int k = j + 1; k < n; k++) A[j * n + k] = 0.0;
  }
}

/* sqrt_info = LLT(cov^-1).matrixL().transpose()   (imu_factor.h:29) */
int orc_imu_sqrt_info(const double *cov, double *sqrt_info) {
  enum { N = 15 };
  double L[N * N], inv[N * N];
  int ok;
  memcpy(L, cov, sizeof L);
  chol_lower(L, N, &ok);
  if (!ok) return 1;
  /* inverse of SPD via its Cholesky factor: solve L L^T X = I column by column */
  for (int c = 0; c < N; c++) {
    double y[N], x[N];
    for (int i = 0; i < N; i++) {
      double s = (i == c) ? 1.0 : 0.0;
      for (int k = 0; k < i; k++) s -= L[i * N + k] * y[k];
      y[i] = s / L[i * N + i];
    }
    for (int i = N - 1; i >= 0; i--) {
      double s = y[i];
      for (int k = i + 1; k < N; k++) s -= L[k * N + i] * x[k];
      x[i] = s / L[i * N + i];
    }
    for (int i = 0; i < N; i++) inv[i * N + c] = x[i];
  }
  for (int i = 0; i < N; i++)
    for (int j = i + 1; j < N; j++) { double m = 0.5 * (inv[i * N + j] + inv[j * N + i]); inv[i * N + j] = inv[j * N + i] = m; }
  chol_lower(inv, N, &ok);
  if (!ok) return 2;
  for (int i = 0; i < N; i++)
    for (int j = 0; j < N; j++) sqrt_info[i * N + j] = inv[j * N + i]; /* transpose of L */
  return 0;
}

#define JB(M, r, c) (&(M)[(r) * 15 + (c)]) /* block pointer in 15x15 row-major */
static void get33(double *o, const double *M, int r, int c) {
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) o[i * 3 + j] = M[(r + i) * 15 + c + j];
}

/* IntegrationBase::evaluate, d2common/include/d2common/integration_base.h:201-227 */
void orc_imu_residual(const orc_imu_const *p, double g_norm, const double *pose_i, const double *sb_i,
                      const double *pose_j, const double *sb_j, double *res15) {
  const double *Pi = pose_i, *Pj = pose_j, *Vi = sb_i, *Bai = sb_i + 3, *Bgi = sb_i + 6;
  const double *Vj = sb_j, *Baj = sb_j + 3, *Bgj = sb_j + 6;
  oq_t Qi = q_from(pose_i + 3), Qj = q_from(pose_j + 3);
  double dp_dba[9], dp_dbg[9], dq_dbg[9], dv_dba[9], dv_dbg[9];
  get33(dp_dba, p->jacobian, 0, 9); get33(dp_dbg, p->jacobian, 0, 12);
  get33(dq_dbg, p->jacobian, 3, 12);
  get33(dv_dba, p->jacobian, 6, 9); get33(dv_dbg, p->jacobian, 6, 12);
  double dba[3], dbg[3], t[3], t2[3];
  v3_sub(dba, Bai, p->linearized_ba); v3_sub(dbg, Bgi, p->linearized_bg);
  m3_vec(t, dq_dbg, dbg);
  oq_t cq = q_mul(q_from(p->delta_q), q_delta(t));
  double cv[3], cp[3];
  m3_vec(t, dv_dba, dba); m3_vec(t2, dv_dbg, dbg);
  for (int k = 0; k < 3; k++) cv[k] = p->delta_v[k] + t[k] + t2[k];
  m3_vec(t, dp_dba, dba); m3_vec(t2, dp_dbg, dbg);
  for (int k = 0; k < 3; k++) cp[k] = p->delta_p[k] + t[k] + t2[k];
  double G[3] = {0, 0, g_norm}, dt = p->sum_dt, a[3];
  for (int k = 0; k < 3; k++) a[k] = 0.5 * G[k] * dt * dt + Pj[k] - Pi[k] - Vi[k] * dt;
  q_rot(t, q_inv(Qi), a);
  for (int k = 0; k < 3; k++) res15[k] = t[k] - cp[k];
  oq_t e = q_mul(q_inv(cq), q_mul(q_inv(Qi), Qj));
  res15[3] = 2 * e.x; res15[4] = 2 * e.y; res15[5] = 2 * e.z;
  for (int k = 0; k < 3; k++) a[k] = G[k] * dt + Vj[k] - Vi[k];
  q_rot(t, q_inv(Qi), a);
  for (int k = 0; k < 3; k++) res15[6 + k] = t[k] - cv[k];
  for (int k = 0; k < 3; k++) { res15[9 + k] = Baj[k] - Bai[k]; res15[12 + k] = Bgj[k] - Bgi[k]; }
}

static void set33(double *J, int ld, int r, int c, const double *m, double s) {
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) J[(r + i) * ld + c + j] = s * m[i * 3 + j];
}
static void left_mul15(double *J, int cols, const double *A) { /* J = A(15x15) * J(15 x cols) */
  double *T = (double *)malloc(sizeof(double) * 15 * cols);
  for (int i = 0; i < 15; i++)
    for (int j = 0; j < cols; j++) {
      double s = 0;
      for (int k = 0; k < 15; k++) s += A[i * 15 + k] * J[k * cols + j];
      T[i * cols + j] = s;
    }
  memcpy(J, T, sizeof(double) * 15 * cols);
  free(T);
}

/* IMUFactor::Evaluate, d2vins/src/factors/imu_factor.h:41-213.
 * Outputs (any Jacobian may be NULL): r(15), J_pose_i 15x7, J_sb_i 15x9, J_pose_j 15x7, J_sb_j 15x9. */
void orc_imu_eval(const orc_imu_const *p, double g_norm, const double *pose_i, const double *sb_i,
                  const double *pose_j, const double *sb_j, double *r, double *J_pose_i, double *J_sb_i,
                  double *J_pose_j, double *J_sb_j) {
  double res[15];
  orc_imu_residual(p, g_norm, pose_i, sb_i, pose_j, sb_j, res);
  for (int i = 0; i < 15; i++) {
    double s = 0;
    for (int k = 0; k < 15; k++) s += p->sqrt_info[i * 15 + k] * res[k];
    r[i] = s;
  }
  if (!J_pose_i && !J_sb_i && !J_pose_j && !J_sb_j) return;
  const double *Pi = pose_i, *Pj = pose_j, *Vi = sb_i, *Bgi = sb_i + 6, *Vj = sb_j;
  oq_t Qi = q_from(pose_i + 3), Qj = q_from(pose_j + 3);
  double dt = p->sum_dt, G[3] = {0, 0, g_norm};
  double dp_dba[9], dp_dbg[9], dq_dbg[9], dv_dba[9], dv_dbg[9];
  get33(dp_dba, p->jacobian, 0, 9); get33(dp_dbg, p->jacobian, 0, 12);
  get33(dq_dbg, p->jacobian, 3, 12);
  get33(dv_dba, p->jacobian, 6, 9); get33(dv_dbg, p->jacobian, 6, 12);
  double RiT[9], t[3], a[3], S[9], M4[16], M4b[16], M4c[16], B3[9];
  q_to_R(RiT, q_inv(Qi)); /* Qi.inverse().toRotationMatrix() */
  double dbg[3];
  v3_sub(dbg, Bgi, p->linearized_bg);
  m3_vec(t, dq_dbg, dbg);
  oq_t dq = q_from(p->delta_q);
  oq_t cq = q_mul(dq, q_delta(t));
  if (J_pose_i) {
    memset(J_pose_i, 0, sizeof(double) * 15 * 7);
    set33(J_pose_i, 7, 0, 0, RiT, -1.0);
    for (int k = 0; k < 3; k++) a[k] = 0.5 * G[k] * dt * dt + Pj[k] - Pi[k] - Vi[k] * dt;
    q_rot(t, q_inv(Qi), a); m3_skew(S, t);
    set33(J_pose_i, 7, 0, 3, S, 1.0);
    q_left(M4, q_mul(q_inv(Qj), Qi)); q_right(M4b, cq);
    mm(M4c, M4, M4b, 4, 4, 4); m4_br3(B3, M4c);
    set33(J_pose_i, 7, 3, 3, B3, -1.0);
    for (int k = 0; k < 3; k++) a[k] = G[k] * dt + Vj[k] - Vi[k];
    q_rot(t, q_inv(Qi), a); m3_skew(S, t);
    set33(J_pose_i, 7, 6, 3, S, 1.0);
    left_mul15(J_pose_i, 7, p->sqrt_info);
  }
  if (J_sb_i) {
    memset(J_sb_i, 0, sizeof(double) * 15 * 9);
    set33(J_sb_i, 9, 0, 0, RiT, -dt);
    set33(J_sb_i, 9, 0, 3, dp_dba, -1.0);
    set33(J_sb_i, 9, 0, 6, dp_dbg, -1.0);
    /* uses the UNcorrected delta_q (imu_factor.h:159) */
    q_left(M4, q_mul(q_mul(q_inv(Qj), Qi), dq)); m4_br3(B3, M4);
    double B3d[9]; m3_mul(B3d, B3, dq_dbg);
    set33(J_sb_i, 9, 3, 6, B3d, -1.0);
    set33(J_sb_i, 9, 6, 0, RiT, -1.0);
    set33(J_sb_i, 9, 6, 3, dv_dba, -1.0);
    set33(J_sb_i, 9, 6, 6, dv_dbg, -1.0);
    double I3[9]; m3_eye(I3);
    set33(J_sb_i, 9, 9, 3, I3, -1.0);
    set33(J_sb_i, 9, 12, 6, I3, -1.0);
    left_mul15(J_sb_i, 9, p->sqrt_info);
  }
  if (J_pose_j) {
    memset(J_pose_j, 0, sizeof(double) * 15 * 7);
    set33(J_pose_j, 7, 0, 0, RiT, 1.0);
    q_left(M4, q_mul(q_mul(q_inv(cq), q_inv(Qi)), Qj)); m4_br3(B3, M4);
    set33(J_pose_j, 7, 3, 3, B3, 1.0);
    left_mul15(J_pose_j, 7, p->sqrt_info);
  }
  if (J_sb_j) {
    memset(J_sb_j, 0, sizeof(double) * 15 * 9);
    double I3[9]; m3_eye(I3);
    set33(J_sb_j, 9, 6, 0, RiT, 1.0);
    set33(J_sb_j, 9, 9, 3, I3, 1.0);
    set33(J_sb_j, 9, 12, 6, I3, 1.0);
    left_mul15(J_sb_j, 9, p->sqrt_info);
  }
}

/* ------------------------------------------------------------------ midpoint pre-integration
 * IntegrationBase::midPointIntegration / propagate, integration_base.h:95-199.
 * noise = diag(acc_n^2, gyr_n^2, acc_n^2, gyr_n^2, acc_w^2, gyr_w^2) (x) I3
 * (d2vins/src/d2vins_params.cpp:58-71). */
void orc_preintegrate(int n, const double *dt, const double *acc /*n+1 x3: acc[0] is acc_0*/,
                      const double *gyr, const double *ba, const double *bg, double acc_n,
                      double gyr_n, double acc_w, double gyr_w, orc_imu_const *out) {
  double noise[18];
  for (int k = 0; k < 3; k++) {
    noise[k] = acc_n * acc_n; noise[3 + k] = gyr_n * gyr_n; noise[6 + k] = acc_n * acc_n;
    noise[9 + k] = gyr_n * gyr_n; noise[12 + k] = acc_w * acc_w; noise[15 + k] = gyr_w * gyr_w;
  }
  double J[225] = {0}, C[225] = {0};
  for (int i = 0; i < 15; i++) J[i * 15 + i] = 1.0;
  double dp[3] = {0, 0, 0}, dv[3] = {0, 0, 0}, sum_dt = 0;
  oq_t dq = {0, 0, 0, 1};
  double acc_0[3], gyr_0[3];
  v3_cpy(acc_0, acc); v3_cpy(gyr_0, gyr);
  for (int s = 0; s < n; s++) {
    const double *acc_1 = acc + 3 * (s + 1), *gyr_1 = gyr + 3 * (s + 1);
    double _dt = dt[s];
    double a0[3], a1[3], w[3], un_acc_0[3], un_acc_1[3], un_acc[3];
    v3_sub(a0, acc_0, ba); v3_sub(a1, acc_1, ba);
    for (int k = 0; k < 3; k++) w[k] = 0.5 * (gyr_0[k] + gyr_1[k]) - bg[k];
    q_rot(un_acc_0, dq, a0);
    oq_t inc = {w[0] * _dt / 2, w[1] * _dt / 2, w[2] * _dt / 2, 1.0};
    oq_t rq = q_mul(dq, inc);
    q_rot(un_acc_1, rq, a1);
    for (int k = 0; k < 3; k++) un_acc[k] = 0.5 * (un_acc_0[k] + un_acc_1[k]);
    double rp[3], rv[3];
    for (int k = 0; k < 3; k++) {
      rp[k] = dp[k] + dv[k] * _dt + 0.5 * un_acc[k] * _dt * _dt;
      rv[k] = dv[k] + un_acc[k] * _dt;
    }
    /* Jacobian / covariance propagation (integration_base.h:118-166) */
    double Rwx[9], Ra0[9], Ra1[9], Rd[9], Rr[9], I3[9], ImW[9], T1[9], T2[9];
    m3_skew(Rwx, w); m3_skew(Ra0, a0); m3_skew(Ra1, a1);
    q_to_R(Rd, dq); q_to_R(Rr, rq); m3_eye(I3);
    for (int k = 0; k < 9; k++) ImW[k] = I3[k] - Rwx[k] * _dt;
    double F[225] = {0}, V[15 * 18] = {0};
    double RdRa0[9], RrRa1[9], RrRa1ImW[9];
    m3_mul(RdRa0, Rd, Ra0); m3_mul(RrRa1, Rr, Ra1); m3_mul(RrRa1ImW, RrRa1, ImW);
    set33(F, 15, 0, 0, I3, 1.0);
    for (int k = 0; k < 9; k++) T1[k] = -0.25 * RdRa0[k] * _dt * _dt + -0.25 * RrRa1ImW[k] * _dt * _dt;
    set33(F, 15, 0, 3, T1, 1.0);
    set33(F, 15, 0, 6, I3, _dt);
    for (int k = 0; k < 9; k++) T1[k] = -0.25 * (Rd[k] + Rr[k]) * _dt * _dt;
    set33(F, 15, 0, 9, T1, 1.0);
    for (int k = 0; k < 9; k++) T1[k] = -0.25 * RrRa1[k] * _dt * _dt * -_dt;
    set33(F, 15, 0, 12, T1, 1.0);
    set33(F, 15, 3, 3, ImW, 1.0);
    set33(F, 15, 3, 12, I3, -1.0 * _dt);
    for (int k = 0; k < 9; k++) T1[k] = -0.5 * RdRa0[k] * _dt + -0.5 * RrRa1ImW[k] * _dt;
    set33(F, 15, 6, 3, T1, 1.0);
    set33(F, 15, 6, 6, I3, 1.0);
    for (int k = 0; k < 9; k++) T1[k] = -0.5 * (Rd[k] + Rr[k]) * _dt;
    set33(F, 15, 6, 9, T1, 1.0);
    for (int k = 0; k < 9; k++) T1[k] = -0.5 * RrRa1[k] * _dt * -_dt;
    set33(F, 15, 6, 12, T1, 1.0);
    set33(F, 15, 9, 9, I3, 1.0);
    set33(F, 15, 12, 12, I3, 1.0);

    set33(V, 18, 0, 0, Rd, 0.25 * _dt * _dt);
    for (int k = 0; k < 9; k++) T2[k] = 0.25 * -RrRa1[k] * _dt * _dt * 0.5 * _dt;
    set33(V, 18, 0, 3, T2, 1.0);
    set33(V, 18, 0, 6, Rr, 0.25 * _dt * _dt);
    set33(V, 18, 0, 9, T2, 1.0);
    set33(V, 18, 3, 3, I3, 0.5 * _dt);
    set33(V, 18, 3, 9, I3, 0.5 * _dt);
    set33(V, 18, 6, 0, Rd, 0.5 * _dt);
    for (int k = 0; k < 9; k++) T2[k] = 0.5 * -RrRa1[k] * _dt * 0.5 * _dt;
    set33(V, 18, 6, 3, T2, 1.0);
    set33(V, 18, 6, 6, Rr, 0.5 * _dt);
    set33(V, 18, 6, 9, T2, 1.0);
    set33(V, 18, 9, 12, I3, _dt);
    set33(V, 18, 12, 15, I3, _dt);

    double FJ[225], FC[225], FCFt[225];
    mm(FJ, F, J, 15, 15, 15);
    memcpy(J, FJ, sizeof J);
    mm(FC, F, C, 15, 15, 15);
    for (int i = 0; i < 15; i++)
      for (int j = 0; j < 15; j++) {
        double s1 = 0, s2 = 0;
        for (int k = 0; k < 15; k++) s1 += FC[i * 15 + k] * F[j * 15 + k];
        for (int k = 0; k < 18; k++) s2 += V[i * 18 + k] * noise[k] * V[j * 18 + k];
        FCFt[i * 15 + j] = s1 + s2;
      }
    memcpy(C, FCFt, sizeof C);
    v3_cpy(dp, rp); v3_cpy(dv, rv);
    dq = q_normalized(rq);
    sum_dt += _dt;
    v3_cpy(acc_0, acc_1); v3_cpy(gyr_0, gyr_1);
  }
  out->sum_dt = sum_dt;
  v3_cpy(out->delta_p, dp); v3_cpy(out->delta_v, dv); q_to(out->delta_q, dq);
  v3_cpy(out->linearized_ba, ba); v3_cpy(out->linearized_bg, bg);
  memcpy(out->jacobian, J, sizeof J);
  memcpy(out->covariance, C, sizeof C);
  orc_imu_sqrt_info(C, out->sqrt_info);
}

/* ------------------------------------------------------------------ consensus factor
 * ConsenusPoseFactor, d2common/src/solver/consenus_factor.cpp:6-51.
 * NB the constructor assigns q_sqrt_info = rho_T * I and T_sqrt_info = rho_theta * I
 * (cpp:15-16): names swapped, weights are rho (not sqrt(rho)). Reproduced. */
void orc_consensus_eval(const double *t_ref, const double *q_ref_xyzw, const double *t_tilde,
                        const double *theta_tilde, double rho_T, double rho_theta, const double *pose,
                        double *r6, double *J6x7) {
  double q_sqrt_info = rho_T, T_sqrt_info = rho_theta;
  oq_t qref = q_from(q_ref_xyzw), ql = q_from(pose + 3);
  double Rref[9], Rinv[9], d[3], t[3];
  q_to_R(Rref, qref); m3_T(Rinv, Rref);
  oq_t qerr = q_mul(q_inv(qref), ql);
  r6[3] = q_sqrt_info * (2.0 * qerr.x + theta_tilde[0]);
  r6[4] = q_sqrt_info * (2.0 * qerr.y + theta_tilde[1]);
  r6[5] = q_sqrt_info * (2.0 * qerr.z + theta_tilde[2]);
  v3_sub(d, pose, t_ref); m3_vec(t, Rinv, d);
  for (int k = 0; k < 3; k++) r6[k] = T_sqrt_info * (t[k] + t_tilde[k]);
  if (J6x7) {
    memset(J6x7, 0, sizeof(double) * 42);
    double M4[16], B3[9];
    q_left(M4, qerr); m4_br3(B3, M4);
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) {
        J6x7[i * 7 + j] = T_sqrt_info * Rinv[i * 3 + j];
        J6x7[(3 + i) * 7 + 3 + j] = q_sqrt_info * B3[i * 3 + j];
      }
  }
}

/* ------------------------------------------------------------------ robust loss
 * ResidualInfo::Evaluate loss section == ceres Corrector,
 * d2common/src/solver/BaseParamResInfo.cpp:71-92, with ceres::HuberLoss(a):
 * rho = s (s<=a^2) else 2a sqrt(s) - a^2; rho' = 1 or a/sqrt(s); rho'' = 0 or -a/(2 s^1.5). */
void orc_huber(double a, double s, double rho[3]) {
  double b = a * a;
  if (s > b) {
    double rr = sqrt(s);
    rho[0] = 2.0 * a * rr - b;
    rho[1] = (a / rr > 2.2250738585072014e-308) ? a / rr : 2.2250738585072014e-308;
    rho[2] = -rho[1] / (2.0 * s);
  } else { rho[0] = s; rho[1] = 1.0; rho[2] = 0.0; }
}
/* returns residual scaling; J scaling = sqrt(rho1) (alpha branch never taken: rho2 <= 0) */
void orc_corrector(const double rho[3], double sq_norm, double *residual_scaling, double *sqrt_rho1,
                   double *alpha_sq_norm) {
  *sqrt_rho1 = sqrt(rho[1]);
  if (sq_norm == 0.0 || rho[2] <= 0.0) { *residual_scaling = *sqrt_rho1; *alpha_sq_norm = 0.0; return; }
  double D = 1.0 + 2.0 * sq_norm * rho[2] / rho[1], alpha = 1.0 - sqrt(D);
  *residual_scaling = *sqrt_rho1 / (1 - alpha);
  *alpha_sq_norm = alpha / sq_norm;
}

/* ------------------------------------------------------------------ manifold
 * PoseLocalParameterization::Plus, d2common/src/solver/pose_local_parameterization.cpp:13-30 */
void orc_pose_plus(const double *x, const double *delta, double *out) {
  out[0] = x[0] + delta[0]; out[1] = x[1] + delta[1]; out[2] = x[2] + delta[2];
  oq_t q = q_normalized(q_mul(q_from(x + 3), q_delta(delta + 3)));
  q_to(out + 3, q);
}

/* ------------------------------------------------------------------ prior dx
 * PriorFactor::Evaluate dx section, d2vins/src/factors/prior_factor.cpp:57-68 */
void orc_prior_dx_pose(const double *x, const double *x0, double *dx6) {
  dx6[0] = x[0] - x0[0]; dx6[1] = x[1] - x0[1]; dx6[2] = x[2] - x0[2];
  oq_t qerr = q_mul(q_inv(q_from(x0 + 3)), q_from(x + 3));
  oq_t p = q_positify(qerr);
  dx6[3] = 2.0 * p.x; dx6[4] = 2.0 * p.y; dx6[5] = 2.0 * p.z;
  if (!(qerr.w >= 0)) { dx6[3] = 2.0 * -p.x; dx6[4] = 2.0 * -p.y; dx6[5] = 2.0 * -p.z; }
}

/* ------------------------------------------------------------------ symmetric eigen (cyclic Jacobi)
 * stands in for Eigen::SelfAdjointEigenSolver; eigenvalues ascending, V columns eigenvectors */
void orc_sym_eig(int n, const double *A_in, double *evals, double *V) {
  double *A = (double *)malloc(sizeof(double) * n * n);
  memcpy(A, A_in, sizeof(double) * n * n);
  for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) V[i * n + j] = (i == j);
  for (int sweep = 0; sweep < 100; sweep++) {
    double off = 0, diag = 0;
    for (int i = 0; i < n; i++) { diag += A[i * n + i] * A[i * n + i]; for (int j = i + 1; j < n; j++) off += A[i * n + j] * A[i * n + j]; }
    if (off <= 1e-32 * (diag + off) || off == 0.0) break;
    for (int p = 0; p < n; p++)
      for (int q = p + 1; q < n; q++) {
        double apq = A[p * n + q];
        if (apq == 0.0) continue;
        double app = A[p * n + p], aqq = A[q * n + q];
        double tau = (aqq - app) / (2.0 * apq);
        double t = (tau >= 0 ? 1.0 : -1.0) / (fabs(tau) + sqrt(1.0 + tau * tau));
        double c = 1.0 / sqrt(1.0 + t * t), s = t * c;
        for (int k = 0; k < n; k++) {
          double akp = A[k * n + p], akq = A[k * n + q];
          A[k * n + p] = c * akp - s * akq; A[k * n + q] = s * akp + c * akq;
        }
        for (int k = 0; k < n; k++) {
          double apk = A[p * n + k], aqk = A[q * n + k];
          A[p * n + k] = c * apk - s * aqk; A[q * n + k] = s * apk + c * aqk;
        }
        for (int k = 0; k < n; k++) {
          double vkp = V[k * n + p], vkq = V[k * n + q];
          V[k * n + p] = c * vkp - s * vkq; V[k * n + q] = s * vkp + c * vkq;
        }
      }
  }
  for (int i = 0; i < n; i++) evals[i] = A[i * n + i];
  /* sort ascending */
  for (int i = 0; i < n; i++) {
    int m = i;
    for (int j = i + 1; j < n; j++) if (evals[j] < evals[m]) m = j;
    if (m != i) {
      double t = evals[i]; evals[i] = evals[m]; evals[m] = t;
      for (int k = 0; k < n; k++) { double v = V[k * n + i]; V[k * n + i] = V[k * n + m]; V[k * n + m] = v; }
    }
  }
  free(A);
}

/* toJacRes, d2vins/src/factors/prior_factor.cpp:132-177: A=(A+A^T)/2, eig, clamp <=1e-8 -> 0,
 * J = sqrt(S) V^T, e0 = sqrt(S^-1) V^T b. */
void orc_to_jac_res(int m, const double *A_in, const double *b, double *J, double *e0) {
  double *A = (double *)malloc(sizeof(double) * m * m), *ev = (double *)malloc(sizeof(double) * m),
         *V = (double *)malloc(sizeof(double) * m * m);
  for (int i = 0; i < m; i++) for (int j = 0; j < m; j++) A[i * m + j] = (A_in[i * m + j] + A_in[j * m + i]) / 2;
  orc_sym_eig(m, A, ev, V);
  const double eps = 1e-8;
  for (int i = 0; i < m; i++) {
    double S = ev[i] > eps ? ev[i] : 0.0, Sinv = ev[i] > eps ? 1.0 / ev[i] : 0.0;
    double ss = sqrt(S), si = sqrt(Sinv), dot = 0;
    for (int k = 0; k < m; k++) { J[i * m + k] = ss * V[k * m + i]; dot += V[k * m + i] * b[k]; }
    e0[i] = si * dot;
  }
  free(A); free(ev); free(V);
}

/* Utility::averageQuaterions, d2common/include/d2common/utils.hpp:213-228: principal eigenvector
 * of sum q q^T over coeffs (x,y,z,w).  A single quaternion is returned unchanged. */
void orc_average_quats(int n, const double *q_xyzw, double *out_xyzw) {
  if (n == 1) { memcpy(out_xyzw, q_xyzw, 4 * sizeof(double)); return; }
  double M[16] = {0}, ev[4], V[16];
  for (int i = 0; i < n; i++)
    for (int a = 0; a < 4; a++) for (int b = 0; b < 4; b++) M[a * 4 + b] += q_xyzw[i * 4 + a] * q_xyzw[i * 4 + b];
  orc_sym_eig(4, M, ev, V);
  for (int a = 0; a < 4; a++) out_xyzw[a] = V[a * 4 + 3];
}

/* Swarm::Pose::DeltaPose(a,b).tangentSpace() -- swarm_msgs is NOT in the reference tree
 * (un-vendored).  ASSUMED semantics (upstream swarm_msgs/Pose.h): DeltaPose = a^-1 * b,
 * tangentSpace = [translation ; angle * axis] with Eigen::AngleAxisd(q) conventions
 * (angle = 2*atan2(|v|, |w|), axis sign follows w). */
void orc_delta_pose_tangent(const double *a, const double *b, double *out6) {
  oq_t qa = q_from(a + 3), qb = q_from(b + 3);
  double d[3], t[3];
  v3_sub(d, b, a);
  q_rot(t, q_inv(qa), d);
  v3_cpy(out6, t);
  oq_t q = q_mul(q_inv(qa), qb);
  double n = sqrt(q.x * q.x + q.y * q.y + q.z * q.z);
  if (n > 0) {
    double angle = 2.0 * atan2(n, fabs(q.w));
    double sgn = q.w < 0 ? -1.0 : 1.0;
    out6[3] = angle * sgn * q.x / n; out6[4] = angle * sgn * q.y / n; out6[5] = angle * sgn * q.z / n;
  } else { out6[3] = out6[4] = out6[5] = 0.0; }
}
