/*
 * orc_math.h -- tiny fixed-size linear algebra for the CPU oracle (TEST INFRASTRUCTURE ONLY).
 *
 * The reference uses Eigen (absent in this container); these helpers restate
 * the few operations its hot path needs.  Quaternions are stored [x y z w] like
 * Eigen::Quaterniond::coeffs() and the reference's parameter blocks.
 *
 * Follows: d2common/include/d2common/utils.hpp:25-104 (deltaQ, positify, skewSymmetric,
 * Qleft, Qright).
 */
#ifndef ORC_MATH_H_
#define ORC_MATH_H_
#include <math.h>
#include <string.h>

typedef struct { double x, y, z, w; } oq_t;

static inline void v3_set(double *o, double a, double b, double c) { o[0] = a; o[1] = b; o[2] = c; }
static inline void v3_cpy(double *o, const double *a) { o[0] = a[0]; o[1] = a[1]; o[2] = a[2]; }
static inline void v3_add(double *o, const double *a, const double *b) { o[0] = a[0] + b[0]; o[1] = a[1] + b[1]; o[2] = a[2] + b[2]; }
static inline void v3_sub(double *o, const double *a, const double *b) { o[0] = a[0] - b[0]; o[1] = a[1] - b[1]; o[2] = a[2] - b[2]; }
static inline void v3_scale(double *o, const double *a, double s) { o[0] = a[0] * s; o[1] = a[1] * s; o[2] = a[2] * s; }
static inline double v3_dot(const double *a, const double *b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
static inline double v3_norm(const double *a) { return sqrt(v3_dot(a, a)); }
static inline void v3_cross(double *o, const double *a, const double *b) {
  double t0 = a[1] * b[2] - a[2] * b[1], t1 = a[2] * b[0] - a[0] * b[2], t2 = a[0] * b[1] - a[1] * b[0];
  o[0] = t0; o[1] = t1; o[2] = t2;
}
/* 3x3 row-major */
static inline void m3_mul(double *o, const double *a, const double *b) {
  double t[9];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) t[i * 3 + j] = a[i * 3] * b[j] + a[i * 3 + 1] * b[3 + j] + a[i * 3 + 2] * b[6 + j];
  memcpy(o, t, sizeof t);
}
static inline void m3_T(double *o, const double *a) {
  double t[9] = {a[0], a[3], a[6], a[1], a[4], a[7], a[2], a[5], a[8]};
  memcpy(o, t, sizeof t);
}
static inline void m3_vec(double *o, const double *a, const double *v) {
  double t0 = a[0] * v[0] + a[1] * v[1] + a[2] * v[2];
  double t1 = a[3] * v[0] + a[4] * v[1] + a[5] * v[2];
  double t2 = a[6] * v[0] + a[7] * v[1] + a[8] * v[2];
  o[0] = t0; o[1] = t1; o[2] = t2;
}
static inline void m3_scale(double *o, const double *a, double s) { for (int i = 0; i < 9; i++) o[i] = a[i] * s; }
static inline void m3_addm(double *o, const double *a, const double *b) { for (int i = 0; i < 9; i++) o[i] = a[i] + b[i]; }
static inline void m3_subm(double *o, const double *a, const double *b) { for (int i = 0; i < 9; i++) o[i] = a[i] - b[i]; }
static inline void m3_eye(double *o) { memset(o, 0, 9 * sizeof(double)); o[0] = o[4] = o[8] = 1.0; }
/* utils.hpp:65-73 */
static inline void m3_skew(double *o, const double *q) {
  o[0] = 0; o[1] = -q[2]; o[2] = q[1];
  o[3] = q[2]; o[4] = 0; o[5] = -q[0];
  o[6] = -q[1]; o[7] = q[0]; o[8] = 0;
}
/* general small dense: o(r x c) = a(r x k) * b(k x c), row-major */
static inline void mm(double *o, const double *a, const double *b, int r, int k, int c) {
  for (int i = 0; i < r; i++)
    for (int j = 0; j < c; j++) {
      double s = 0;
      for (int t = 0; t < k; t++) s += a[i * k + t] * b[t * c + j];
      o[i * c + j] = s;
    }
}

/* ---- quaternions (Hamilton, [x y z w]) ---- */
static inline oq_t q_from(const double *p) { oq_t q = {p[0], p[1], p[2], p[3]}; return q; }
static inline void q_to(double *p, oq_t q) { p[0] = q.x; p[1] = q.y; p[2] = q.z; p[3] = q.w; }
static inline oq_t q_mul(oq_t a, oq_t b) {
  oq_t r;
  r.w = a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z;
  r.x = a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y;
  r.y = a.w * b.y - a.x * b.z + a.y * b.w + a.z * b.x;
  r.z = a.w * b.z + a.x * b.y - a.y * b.x + a.z * b.w;
  return r;
}
/* Eigen::Quaternion::inverse(): conjugate / squaredNorm */
static inline oq_t q_inv(oq_t a) {
  double n2 = a.x * a.x + a.y * a.y + a.z * a.z + a.w * a.w;
  oq_t r = {-a.x / n2, -a.y / n2, -a.z / n2, a.w / n2};
  return r;
}
static inline oq_t q_normalized(oq_t a) {
  double n = sqrt(a.x * a.x + a.y * a.y + a.z * a.z + a.w * a.w);
  oq_t r = {a.x / n, a.y / n, a.z / n, a.w / n};
  return r;
}
/* Eigen::Quaternion::toRotationMatrix() (no normalisation, same arithmetic form) */
static inline void q_to_R(double *R, oq_t q) {
  double tx = 2 * q.x, ty = 2 * q.y, tz = 2 * q.z;
  double twx = tx * q.w, twy = ty * q.w, twz = tz * q.w;
  double txx = tx * q.x, txy = ty * q.x, txz = tz * q.x;
  double tyy = ty * q.y, tyz = tz * q.y, tzz = tz * q.z;
  R[0] = 1 - (tyy + tzz); R[1] = txy - twz; R[2] = txz + twy;
  R[3] = txy + twz; R[4] = 1 - (txx + tzz); R[5] = tyz - twx;
  R[6] = txz - twy; R[7] = tyz + twx; R[8] = 1 - (txx + tyy);
}
/* Eigen: q * v  (rotation of a vector; uses the uv / uuv form) */
static inline void q_rot(double *o, oq_t q, const double *v) {
  double u[3] = {q.x, q.y, q.z}, uv[3], uuv[3];
  v3_cross(uv, u, v);
  uv[0] += uv[0]; uv[1] += uv[1]; uv[2] += uv[2];
  v3_cross(uuv, u, uv);
  o[0] = v[0] + q.w * uv[0] + uuv[0];
  o[1] = v[1] + q.w * uv[1] + uuv[1];
  o[2] = v[2] + q.w * uv[2] + uuv[2];
}
/* utils.hpp:25-38 */
static inline oq_t q_delta(const double *theta) { oq_t r = {theta[0] / 2, theta[1] / 2, theta[2] / 2, 1.0}; return r; }
/* utils.hpp:56-63 */
static inline oq_t q_positify(oq_t q) {
  if (q.w >= 0.0) return q;
  oq_t p = {-q.x, -q.y, -q.z, -q.w};
  return p;
}
/* utils.hpp:85-104, 4x4 row-major with (w, x, y, z) ordering */
static inline void q_left(double *M, oq_t q0) {
  oq_t q = q_positify(q0);
  double v[3] = {q.x, q.y, q.z}, S[9];
  m3_skew(S, v);
  M[0] = q.w; M[1] = -q.x; M[2] = -q.y; M[3] = -q.z;
  for (int i = 0; i < 3; i++) {
    M[(i + 1) * 4] = v[i];
    for (int j = 0; j < 3; j++) M[(i + 1) * 4 + 1 + j] = (i == j ? q.w : 0.0) + S[i * 3 + j];
  }
}
static inline void q_right(double *M, oq_t p0) {
  oq_t p = q_positify(p0);
  double v[3] = {p.x, p.y, p.z}, S[9];
  m3_skew(S, v);
  M[0] = p.w; M[1] = -p.x; M[2] = -p.y; M[3] = -p.z;
  for (int i = 0; i < 3; i++) {
    M[(i + 1) * 4] = v[i];
    for (int j = 0; j < 3; j++) M[(i + 1) * 4 + 1 + j] = (i == j ? p.w : 0.0) - S[i * 3 + j];
  }
}
static inline void m4_br3(double *o, const double *M) { /* bottomRightCorner<3,3>() */
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) o[i * 3 + j] = M[(i + 1) * 4 + 1 + j];
}
#endif
