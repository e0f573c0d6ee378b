// d2pgo.cu -- pose-graph optimisation on the GPU (include/d2pgo.h): relative-pose factors, matrix-free block-Jacobi PCG,
// Levenberg-Marquardt outer loop, edge-sharded multi-GPU with one NCCL all-reduce per CG iteration.
//
// Factor: D2Common::RelPoseFactorAD (d2common/include/d2common/solver/RelPoseFactor.hpp:68-135), restated in pgo_edge_eval
// below with analytic exact derivatives in the tangent of the right-multiplicative pose retraction
// (pose_local_parameterization.cpp:13-38).  (The header's hand-differentiated RelPoseFactor :8-66, used when
// pgo_use_autodiff is off, drops q_rel from its rotation blocks -- its Jacobian is exact only for identity relative rotation.)
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <string.h>

#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/d2pgo.h"
#include "d2ba_math.cuh"

namespace d2ba {
// NCCL through the dlopen'ed entry points of d2ba_host.cu
int nccl_comm_init(void **comm, const uint8_t *unique_id, int rank, int nranks, std::string &err);
int nccl_allreduce_f64(void *comm, double *buf, size_t n, cudaStream_t s);
void nccl_comm_destroy(void *comm);
}  // namespace d2ba
using namespace d2ba;

namespace {

struct PgoScalars {      // device-resident CG state that is not a per-block partial
  double bb;             // |b|^2 of the current linear system
  int done, iters;
};

struct PgoDev {
  int n_pose, n_edge;
  const double *x;        // [N][8] poses the edges are linearised at
  const unsigned char *fixed;
  const int *ea, *eb;     // [E] pose indices
  const double *rel;      // [E][8]
  const double *sinfo;    // [E][36] sqrt information, row-major
  double *lin;            // [78][E] field-major: r(6), J0 (6x6 row-major), J1 (6x6)
  double *g, *D;          // [6N], [N][36]
  double *Minv;           // [N][36]
  double *dx, *r, *z, *p, *Ap;   // [6N]
  double *damp;           // [6N] lambda diag(D) + 1e-12
  double *t;              // [12][E] per edge: J0^T (J p) (6), J1^T (J p) (6)
  const int *inc_ptr, *inc;   // incidence lists: pose i -> (edge << 1 | side) of this rank's edges, ascending
  double *rz_part, *rr_part;  // [2][nbp] per-block partials, ping-pong on the iteration parity
  double *pAp_part;           // [nbp]
  int nbp;                    // pose blocks of 128
  PgoScalars *s;
};

// RelPoseFactorAD (RelPoseFactor.hpp:68-135; the reference's default 6-DoF factor, pgo_use_autodiff = true in
// d2pgo/src/d2pgo_config.h:52): r = S [ q_a^-1 (p_b - p_a) - p_meas ; 2 vec(q_meas (q_a^-1 q_b)^-1) ] with the full 6x6 S.
// The reference differentiates it with ceres autodiff on the EigenQuaternionManifold; here the exact derivatives are analytic,
// in the tangent of the right-multiplicative pose retraction (the minimiser does not depend on the manifold chart):
//   d p_ab / d dp_a = -Ra^T,  d p_ab / d dth_a = [p_ab]x,  d p_ab / d dp_b = Ra^T,
//   d 2vec(dq) / d dth_a = (w I + [v]x) of dq,   d 2vec(dq) / d dth_b = -(Qleft(q_meas) Qright(q_b^-1 q_a))_3.
D2BA_DEV void pgo_edge_eval(const double *p0, const double *p1, const double *rel, const double *S, double *r, double *J0, double *J1) {
  const Q4 q0 = qload(p0 + 3), q1 = qload(p1 + 3), qm = qload(rel + 3);
  const Q4 q0i = Q4{-q0.x, -q0.y, -q0.z, q0.w}, q1i = Q4{-q1.x, -q1.y, -q1.z, q1.w};   // conjugates (:83)
  double R0i[9];
  q2R(q0i, R0i);
  const double dt[3] = {p1[0] - p0[0], p1[1] - p0[1], p1[2] - p0[2]};
  double pab[3];
  mv3(R0i, dt, pab);                                  // q_a_inverse * (p_b - p_a) (:87)
  const Q4 X = qmul(q1i, q0);                         // (q_a^-1 q_b)^-1
  const Q4 dq = qmul(qm, X);                          // q_measured * q_ab_estimated.conjugate() (:90-91)
  const double raw[6] = {pab[0] - rel[0], pab[1] - rel[1], pab[2] - rel[2], 2.0 * dq.x, 2.0 * dq.y, 2.0 * dq.z};
  for (int i = 0; i < 6; i++) { double t = 0; for (int k = 0; k < 6; k++) t += S[i * 6 + k] * raw[k]; r[i] = t; }   // applyOnTheLeft (:104)
  if (!J0) return;
  double A0[36], A1[36];
  for (int i = 0; i < 36; i++) { A0[i] = 0.0; A1[i] = 0.0; }
  const double Sk[9] = {0, -pab[2], pab[1], pab[2], 0, -pab[0], -pab[1], pab[0], 0};
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) { A0[i * 6 + j] = -R0i[i * 3 + j]; A1[i * 6 + j] = R0i[i * 3 + j]; A0[i * 6 + 3 + j] = Sk[i * 3 + j]; }
  const double L[9] = {dq.w, -dq.z, dq.y, dq.z, dq.w, -dq.x, -dq.y, dq.x, dq.w};   // w I + [v]x of dq
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) A0[(3 + i) * 6 + 3 + j] = L[i * 3 + j];
  const double Lm[9] = {qm.w, -qm.z, qm.y, qm.z, qm.w, -qm.x, -qm.y, qm.x, qm.w};  // w I + [v]x of q_meas
  const double Rx[9] = {X.w, X.z, -X.y, -X.z, X.w, X.x, X.y, -X.x, X.w};            // w I - [v]x of X
  double P[9];
  mm3(Lm, Rx, P);
  const double vm[3] = {qm.x, qm.y, qm.z}, vx[3] = {X.x, X.y, X.z};
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) A1[(3 + i) * 6 + 3 + j] = -(P[i * 3 + j] - vm[i] * vx[j]);
  for (int i = 0; i < 6; i++) for (int j = 0; j < 6; j++) {
    double t0 = 0, t1 = 0;
    for (int k = 0; k < 6; k++) { t0 += S[i * 6 + k] * A0[k * 6 + j]; t1 += S[i * 6 + k] * A1[k * 6 + j]; }
    J0[i * 6 + j] = t0; J1[i * 6 + j] = t1;
  }
}

// Fixed-order sum of n per-block partials, the same value in every thread of every block (so that all blocks -- and all
// ranks, which hold identical vectors -- take the same branch without a single-thread "decide" kernel in between).
D2BA_DEV double total_of(const double *part, int n, double *red) {
  __syncthreads();
  if (threadIdx.x < 32) {
    double v = 0.0;
    for (int i = threadIdx.x; i < n; i += 32) v += part[i];
    v = warp_sum(v);
    if (threadIdx.x == 0) red[39] = v;
  }
  __syncthreads();
  return red[39];
}

// linearise every local edge at d.x: lin records [r | J0 | J1] and per-block cost partials (summed in fixed order by k_pgo_sum)
__global__ void __launch_bounds__(128) k_pgo_lin(PgoDev d, double *cost_part, int want_jac) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  __shared__ double red[40];
  double c = 0.0;
  if (e < d.n_edge) {
    const int a = d.ea[e], b = d.eb[e];
    double r[6], J0[36], J1[36];
    pgo_edge_eval(d.x + (size_t)a * 8, d.x + (size_t)b * 8, d.rel + (size_t)e * 8, d.sinfo + (size_t)e * 36, r, want_jac ? J0 : nullptr, J1);
    for (int k = 0; k < 6; k++) c += 0.5 * r[k] * r[k];
    if (want_jac) {
      double *o = d.lin + e;   // field-major [78][E]: consecutive edges (threads) touch consecutive addresses
      const size_t E = (size_t)d.n_edge;
      for (int k = 0; k < 6; k++) o[k * E] = r[k];
      for (int k = 0; k < 36; k++) { o[(6 + k) * E] = J0[k]; o[(42 + k) * E] = J1[k]; }
    }
  }
  c = block_sum(c, red);
  if (threadIdx.x == 0) cost_part[blockIdx.x] = c;
}
__global__ void k_pgo_sum(const double *part, int n, double *out) {
  __shared__ double red[40];
  const double v = total_of(part, n, red);
  if (threadIdx.x == 0) out[0] = v;
}

// gradient g_i = sum J^T r and block diagonal D_i = sum J^T J over the edges incident to pose i, in the fixed order of the
// incidence list (no atomics: bitwise reproducible)
__global__ void __launch_bounds__(128) k_pgo_gD(PgoDev d, double *g, double *D) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= d.n_pose) return;
  double gi[6] = {0, 0, 0, 0, 0, 0}, Di[36];
  for (int k = 0; k < 36; k++) Di[k] = 0.0;
  if (!d.fixed[i])
    for (int q = d.inc_ptr[i]; q < d.inc_ptr[i + 1]; q++) {
      const int c = d.inc[q];
      const size_t E = (size_t)d.n_edge;
      const double *o = d.lin + (c >> 1), *J = o + (size_t)(6 + 36 * (c & 1)) * E;
      for (int k = 0; k < 6; k++) {
        const double rk = o[k * E];
        double row[6];
        for (int a = 0; a < 6; a++) { row[a] = J[(size_t)(k * 6 + a) * E]; gi[a] += row[a] * rk; }
        for (int a = 0; a < 6; a++) for (int b = 0; b < 6; b++) Di[a * 6 + b] += row[a] * row[b];
      }
    }
  for (int k = 0; k < 6; k++) g[(size_t)i * 6 + k] = gi[k];
  for (int k = 0; k < 36; k++) D[(size_t)i * 36 + k] = Di[k];
}

// block-Jacobi preconditioner: Minv = (D + lambda diag(D))^-1 per free pose (6x6 Cholesky, one thread per pose); also the
// damping diagonal lambda diag(D) + 1e-12 the products add
__global__ void k_pgo_precond(PgoDev d, const double *D, double lambda) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= d.n_pose) return;
  double *M = d.Minv + (size_t)i * 36;
  if (d.fixed[i]) { for (int k = 0; k < 36; k++) M[k] = 0.0; for (int k = 0; k < 6; k++) d.damp[(size_t)i * 6 + k] = 0.0; return; }
  double A[36], Li[36];
  for (int k = 0; k < 36; k++) A[k] = D[(size_t)i * 36 + k];
  for (int k = 0; k < 6; k++) { const double dk = lambda * A[k * 7] + 1e-12; d.damp[(size_t)i * 6 + k] = dk; A[k * 7] += dk; }
  for (int j = 0; j < 6; j++) {   // Cholesky, lower
    double s = A[j * 6 + j];
    for (int k = 0; k < j; k++) s -= A[j * 6 + k] * A[j * 6 + k];
    s = sqrt(s > 0.0 ? s : 1e-300);
    A[j * 6 + j] = s;
    for (int r = j + 1; r < 6; r++) { double t = A[r * 6 + j]; for (int k = 0; k < j; k++) t -= A[r * 6 + k] * A[j * 6 + k]; A[r * 6 + j] = t / s; }
  }
  for (int c = 0; c < 6; c++) {   // L^-1 column by column
    for (int r = 0; r < 6; r++) {
      double t = r == c ? 1.0 : 0.0;
      for (int k = 0; k < r; k++) t -= A[r * 6 + k] * Li[k * 6 + c];
      Li[r * 6 + c] = t / A[r * 6 + r];
    }
  }
  for (int r = 0; r < 6; r++) for (int c = 0; c < 6; c++) { double t = 0; for (int k = 0; k < 6; k++) t += Li[k * 6 + r] * Li[k * 6 + c]; M[r * 6 + c] = t; }
}

// ---- conjugate gradients.  One iteration = three kernels (the three grid-wide dependencies of CG):
//   k_pgo_cg_edge : per edge J0^T t, J1^T t with t = J0 p_a + J1 p_b, p = z + beta p_old formed on the fly   (needs rz of the last update)
//   k_pgo_cg_pose : stores p, Ap_i = sum over incident edges (+ damping), partials of p.Ap               (needs every edge)
//   k_pgo_cg_step : alpha = rz / pAp; dx += alpha p; r -= alpha Ap; z = Minv r; partials of r.z, r.r   (needs p.Ap)
// Scalars live as per-block partials summed in fixed order by every block (total_of); rz / rr ping-pong on the iteration
// parity `par`.  Once converged the sticky flag s->done turns the remaining launches of a graph into no-ops.
struct CgView { double rz_prev, rz_cur, rr_cur; bool done, first; };
D2BA_DEV CgView cg_view(const PgoDev &d, int par, double tol2, double *red) {
  CgView v;
  const PgoScalars *s = d.s;
  v.first = s->iters == 0;
  v.rz_cur = total_of(d.rz_part + (size_t)(par ^ 1) * d.nbp, d.nbp, red);
  v.rr_cur = total_of(d.rr_part + (size_t)(par ^ 1) * d.nbp, d.nbp, red);
  v.rz_prev = v.first ? 1.0 : total_of(d.rz_part + (size_t)par * d.nbp, d.nbp, red);
  v.done = s->done || !(v.rr_cur > tol2 * s->bb) || !(v.rz_cur > 0.0) || !isfinite(v.rz_cur);
  return v;
}
D2BA_DEV void cg_p_new(const PgoDev &d, int i, double beta, double *p) {
  for (int k = 0; k < 6; k++) p[k] = d.z[(size_t)i * 6 + k] + beta * d.p[(size_t)i * 6 + k];   // p holds 0 before the first iteration
}

__global__ void __launch_bounds__(128) k_pgo_cg_edge(PgoDev d, int par, double tol2) {
  __shared__ double red[40];
  const CgView v = cg_view(d, par, tol2, red);
  if (v.done) return;
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= d.n_edge) return;
  const double beta = v.first ? 0.0 : v.rz_cur / v.rz_prev;
  const int a = d.ea[e], b = d.eb[e];
  const size_t E = (size_t)d.n_edge;
  const double *J0 = d.lin + 6 * E + e, *J1 = J0 + 36 * E;
  double pa[6], pb[6], ya[6] = {0, 0, 0, 0, 0, 0}, yb[6] = {0, 0, 0, 0, 0, 0};
  cg_p_new(d, a, beta, pa); cg_p_new(d, b, beta, pb);
  // t = J0 pa + J1 pb, then this edge's two contributions J0^T t, J1^T t (the Jacobians are read once, coalesced; the
  // per-pose kernel only sums six numbers per incident edge)
  for (int k = 0; k < 6; k++) {
    double j0[6], j1[6], tk = 0;
    for (int q = 0; q < 6; q++) { j0[q] = J0[(size_t)(k * 6 + q) * E]; j1[q] = J1[(size_t)(k * 6 + q) * E]; tk += j0[q] * pa[q] + j1[q] * pb[q]; }
    for (int q = 0; q < 6; q++) { ya[q] += j0[q] * tk; yb[q] += j1[q] * tk; }
  }
  for (int q = 0; q < 6; q++) { d.t[(size_t)q * E + e] = ya[q]; d.t[(size_t)(6 + q) * E + e] = yb[q]; }
}

// LOCAL: this rank's edges only, no damping / p.Ap yet (the all-reduce of Ap comes first; k_pgo_cg_pAp follows)
template <bool LOCAL>
__global__ void __launch_bounds__(128) k_pgo_cg_pose(PgoDev d, int par, double tol2) {
  __shared__ double red[40];
  const CgView v = cg_view(d, par, tol2, red);
  if (v.done) return;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  double s = 0.0;
  if (i < d.n_pose) {
    const double beta = v.first ? 0.0 : v.rz_cur / v.rz_prev;
    double p[6], y[6] = {0, 0, 0, 0, 0, 0};
    cg_p_new(d, i, beta, p);
    if (!d.fixed[i])
      for (int q = d.inc_ptr[i]; q < d.inc_ptr[i + 1]; q++) {
        const int c = d.inc[q];
        const size_t E = (size_t)d.n_edge;
        const double *t = d.t + (size_t)(6 * (c & 1)) * E + (c >> 1);
        for (int a = 0; a < 6; a++) y[a] += t[a * E];
      }
    for (int k = 0; k < 6; k++) {
      d.p[(size_t)i * 6 + k] = p[k];
      if (!LOCAL) { y[k] += d.damp[(size_t)i * 6 + k] * p[k]; s += p[k] * y[k]; }
      d.Ap[(size_t)i * 6 + k] = y[k];
    }
  }
  if (!LOCAL) { s = block_sum(s, red); if (threadIdx.x == 0) d.pAp_part[blockIdx.x] = s; }
}
template __global__ void k_pgo_cg_pose<true>(PgoDev, int, double);
template __global__ void k_pgo_cg_pose<false>(PgoDev, int, double);

__global__ void __launch_bounds__(128) k_pgo_cg_pAp(PgoDev d, int par, double tol2) {   // multi-rank: after the all-reduce of Ap
  __shared__ double red[40];
  const CgView v = cg_view(d, par, tol2, red);
  if (v.done) return;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  double s = 0.0;
  if (i < d.n_pose)
    for (int k = 0; k < 6; k++) {
      const double pk = d.p[(size_t)i * 6 + k], y = d.Ap[(size_t)i * 6 + k] + d.damp[(size_t)i * 6 + k] * pk;
      d.Ap[(size_t)i * 6 + k] = y; s += pk * y;
    }
  s = block_sum(s, red);
  if (threadIdx.x == 0) d.pAp_part[blockIdx.x] = s;
}

__global__ void __launch_bounds__(128) k_pgo_cg_step(PgoDev d, int par, double tol2) {
  __shared__ double red[40];
  const CgView v = cg_view(d, par, tol2, red);
  if (v.done) { if (blockIdx.x == 0 && threadIdx.x == 0) d.s->done = 1; return; }
  const double pAp = total_of(d.pAp_part, d.nbp, red);
  const double alpha = pAp > 0.0 ? v.rz_cur / pAp : 0.0;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  double rz = 0, rr = 0;
  if (i < d.n_pose && !d.fixed[i]) {
    double r[6];
    for (int k = 0; k < 6; k++) {
      d.dx[(size_t)i * 6 + k] += alpha * d.p[(size_t)i * 6 + k];
      r[k] = d.r[(size_t)i * 6 + k] - alpha * d.Ap[(size_t)i * 6 + k];
      d.r[(size_t)i * 6 + k] = r[k];
    }
    const double *M = d.Minv + (size_t)i * 36;
    for (int k = 0; k < 6; k++) { double t = 0; for (int q = 0; q < 6; q++) t += M[k * 6 + q] * r[q]; d.z[(size_t)i * 6 + k] = t; rz += r[k] * t; rr += r[k] * r[k]; }
  }
  rz = block_sum(rz, red); rr = block_sum(rr, red);
  if (threadIdx.x == 0) {
    d.rz_part[(size_t)par * d.nbp + blockIdx.x] = rz; d.rr_part[(size_t)par * d.nbp + blockIdx.x] = rr;
    if (blockIdx.x == 0) d.s->iters = d.s->iters + 1;   // read by the next kernel, not by this one's other blocks (cg_view ran before)
  }
}

// CG start: dx = 0, r = -g, z = Minv r, p = 0 (the first k_pgo_cg_edge forms p = z); partials into the parity-1 slots
__global__ void __launch_bounds__(128) k_pgo_cg_init(PgoDev d, const double *g) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  __shared__ double red[40];
  double rz = 0, rr = 0;
  if (i < d.n_pose) {
    double r[6], z[6];
    for (int k = 0; k < 6; k++) r[k] = d.fixed[i] ? 0.0 : -g[(size_t)i * 6 + k];
    const double *M = d.Minv + (size_t)i * 36;
    for (int k = 0; k < 6; k++) { double t = 0; for (int q = 0; q < 6; q++) t += M[k * 6 + q] * r[q]; z[k] = t; }
    for (int k = 0; k < 6; k++) {
      d.dx[(size_t)i * 6 + k] = 0.0; d.r[(size_t)i * 6 + k] = r[k]; d.z[(size_t)i * 6 + k] = z[k]; d.p[(size_t)i * 6 + k] = 0.0;
      rz += r[k] * z[k]; rr += r[k] * r[k];
    }
  }
  rz = block_sum(rz, red); rr = block_sum(rr, red);
  if (threadIdx.x == 0) { d.rz_part[(size_t)d.nbp + blockIdx.x] = rz; d.rr_part[(size_t)d.nbp + blockIdx.x] = rr; }
}
__global__ void k_pgo_cg_init2(PgoDev d) {   // |b|^2 and the counters
  __shared__ double red[40];
  const double bb = total_of(d.rr_part + d.nbp, d.nbp, red);
  if (threadIdx.x == 0) { d.s->bb = bb; d.s->done = 0; d.s->iters = 0; }
}

// candidate poses: x_out = x (+) dx   (PoseLocalParameterization::Plus)
__global__ void k_pgo_retract(PgoDev d, double *x_out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= d.n_pose) return;
  const double *x = d.x + (size_t)i * 8;
  double o[7];
  if (d.fixed[i]) { for (int k = 0; k < 7; k++) o[k] = x[k]; }
  else pose_plus(x, d.dx + (size_t)i * 6, o);
  for (int k = 0; k < 7; k++) x_out[(size_t)i * 8 + k] = o[k];
  x_out[(size_t)i * 8 + 7] = 0.0;
}

template <typename T> struct Buf {
  T *p = nullptr; size_t n = 0;
  cudaError_t alloc(size_t c) { if (c <= n && p) return cudaSuccess; if (p) cudaFree(p); p = nullptr; n = 0; cudaError_t e = cudaMalloc(&p, (c ? c : 1) * sizeof(T)); if (e == cudaSuccess) n = c; return e; }
  void release() { if (p) cudaFree(p); p = nullptr; n = 0; }
};
}  // namespace

struct d2pgo_handle {
  d2pgo_config cfg;
  std::string err;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  std::vector<int64_t> ids; std::unordered_map<int64_t, int> index;
  std::vector<double> poses; std::vector<unsigned char> fixed;
  std::vector<int> ea, eb; std::vector<double> rel, sinfo;
  Buf<double> d_x[2], d_rel, d_sinfo, d_lin[2], d_g[2], d_D[2], d_Minv, d_dx, d_r, d_z, d_p, d_Ap, d_cost, d_damp, d_t, d_part, d_cost_part;
  Buf<unsigned char> d_fixed; Buf<int> d_ea, d_eb, d_inc_ptr, d_inc; Buf<PgoScalars> d_s;
  cudaGraphExec_t cg_graph[2] = {nullptr, nullptr};   // 16 CG iterations on the linearisation buffer 0 / 1 (single rank)
  double graph_tol2 = -1.0;
  bool uploaded = false;
  void *comm = nullptr; int rank = 0, nranks = 1;
};

#define PCK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { h->err = std::string(#call) + ": " + cudaGetErrorString(e_); return 100 + (int)e_; } } while (0)

extern "C" {

int d2pgo_default_config(d2pgo_config *c) {
  if (!c) return 1;
  memset(c, 0, sizeof *c);
  c->device = 0; c->max_iterations = 20; c->pcg_max_iterations = 200; c->pcg_tolerance = 1e-8; c->lambda0 = 1e-6; c->function_tolerance = 1e-9;
  return 0;
}

int d2pgo_create(const d2pgo_config *cfg, d2pgo_handle **out) {
  if (!cfg || !out) return 1;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) { fprintf(stderr, "d2pgo_create: no CUDA device (there is no CPU fallback)\n"); return 3; }
  if (cfg->device < 0 || cfg->device >= ndev || cudaSetDevice(cfg->device) != cudaSuccess) return 4;
  d2pgo_handle *h = new d2pgo_handle();
  h->cfg = *cfg;
  if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) { delete h; return 6; }
  cudaEventCreate(&h->ev0); cudaEventCreate(&h->ev1);
  *out = h;
  return 0;
}

int d2pgo_destroy(d2pgo_handle *h) {
  if (!h) return 0;
  cudaSetDevice(h->cfg.device);
  cudaStreamSynchronize(h->stream);
  if (h->comm) nccl_comm_destroy(h->comm);
  for (int b = 0; b < 2; b++) { h->d_x[b].release(); h->d_g[b].release(); h->d_D[b].release(); h->d_lin[b].release(); if (h->cg_graph[b]) cudaGraphExecDestroy(h->cg_graph[b]); }
  h->d_damp.release(); h->d_t.release(); h->d_part.release(); h->d_cost_part.release(); h->d_inc_ptr.release(); h->d_inc.release();
  h->d_rel.release(); h->d_sinfo.release(); h->d_Minv.release(); h->d_dx.release(); h->d_r.release(); h->d_z.release(); h->d_p.release();
  h->d_Ap.release(); h->d_cost.release(); h->d_fixed.release(); h->d_ea.release(); h->d_eb.release(); h->d_s.release();
  cudaEventDestroy(h->ev0); cudaEventDestroy(h->ev1); cudaStreamDestroy(h->stream);
  delete h;
  return 0;
}

const char *d2pgo_last_error(const d2pgo_handle *h) { return h ? h->err.c_str() : "null handle"; }

int d2pgo_set_poses(d2pgo_handle *h, int32_t n, const int64_t *ids, const double *poses7, const uint8_t *fixed) {
  if (!h || n <= 0) return 1;
  h->ids.assign(ids, ids + n); h->index.clear(); h->poses.assign((size_t)n * 8, 0.0); h->fixed.assign(n, 0);
  for (int i = 0; i < n; i++) {
    if (!h->index.emplace(ids[i], i).second) { h->err = "set_poses: duplicate pose id"; return 2; }
    memcpy(&h->poses[(size_t)i * 8], poses7 + (size_t)i * 7, 56);
    h->fixed[i] = fixed ? fixed[i] : 0;
  }
  h->ea.clear(); h->eb.clear(); h->rel.clear(); h->sinfo.clear(); h->uploaded = false;
  return 0;
}

int d2pgo_add_edges(d2pgo_handle *h, int32_t n, const int64_t *id_a, const int64_t *id_b, const double *rel7, const double *sqrt_info36) {
  if (!h) return 1;
  for (int e = 0; e < n; e++) {
    auto a = h->index.find(id_a[e]), b = h->index.find(id_b[e]);
    if (a == h->index.end() || b == h->index.end()) { h->err = "add_edges: unknown pose id"; return 2; }
    h->ea.push_back(a->second); h->eb.push_back(b->second);
    for (int k = 0; k < 7; k++) h->rel.push_back(rel7[(size_t)e * 7 + k]);
    h->rel.push_back(0.0);
    for (int k = 0; k < 36; k++) h->sinfo.push_back(sqrt_info36[(size_t)e * 36 + k]);
  }
  h->uploaded = false;
  return 0;
}

int d2pgo_comm_init(d2pgo_handle *h, const uint8_t unique_id[128], int32_t rank, int32_t nranks) {
  if (!h) return 1;
  cudaSetDevice(h->cfg.device);
  if (nccl_comm_init(&h->comm, unique_id, rank, nranks, h->err)) return 3;
  h->rank = rank; h->nranks = nranks;
  return 0;
}

static void pgo_drop_graphs(d2pgo_handle *h) {
  for (int b = 0; b < 2; b++) if (h->cg_graph[b]) { cudaGraphExecDestroy(h->cg_graph[b]); h->cg_graph[b] = nullptr; }
}

static int pgo_upload(d2pgo_handle *h) {
  const size_t N = h->ids.size(), E = h->ea.size(), nbp = (N + 127) / 128, nbe = (E + 127) / 128;
  pgo_drop_graphs(h);
  for (int b = 0; b < 2; b++) { PCK(h->d_x[b].alloc(N * 8)); PCK(h->d_g[b].alloc(N * 6)); PCK(h->d_D[b].alloc(N * 36)); PCK(h->d_lin[b].alloc(E * 78)); }
  PCK(h->d_rel.alloc(E * 8)); PCK(h->d_sinfo.alloc(E * 36)); PCK(h->d_Minv.alloc(N * 36));
  PCK(h->d_dx.alloc(N * 6)); PCK(h->d_r.alloc(N * 6)); PCK(h->d_z.alloc(N * 6)); PCK(h->d_p.alloc(N * 6)); PCK(h->d_Ap.alloc(N * 6)); PCK(h->d_cost.alloc(2));
  PCK(h->d_damp.alloc(N * 6)); PCK(h->d_t.alloc(E * 12)); PCK(h->d_part.alloc(5 * nbp)); PCK(h->d_cost_part.alloc(nbe + 1));
  PCK(h->d_fixed.alloc(N)); PCK(h->d_ea.alloc(E)); PCK(h->d_eb.alloc(E)); PCK(h->d_s.alloc(1)); PCK(h->d_inc_ptr.alloc(N + 1)); PCK(h->d_inc.alloc(2 * E));
  // incidence lists (pose -> its edges, ascending edge order): the fixed summation order of every product
  std::vector<int> ptr(N + 1, 0), inc(2 * E);
  for (size_t e = 0; e < E; e++) { ptr[h->ea[e] + 1]++; ptr[h->eb[e] + 1]++; }
  for (size_t i = 0; i < N; i++) ptr[i + 1] += ptr[i];
  { std::vector<int> fill(ptr.begin(), ptr.end() - 1);
    for (size_t e = 0; e < E; e++) { inc[fill[h->ea[e]]++] = (int)(e << 1); inc[fill[h->eb[e]]++] = (int)(e << 1 | 1); } }
  PCK(cudaMemcpyAsync(h->d_x[0].p, h->poses.data(), N * 64, cudaMemcpyHostToDevice, h->stream));
  PCK(cudaMemcpyAsync(h->d_fixed.p, h->fixed.data(), N, cudaMemcpyHostToDevice, h->stream));
  PCK(cudaMemcpyAsync(h->d_inc_ptr.p, ptr.data(), (N + 1) * 4, cudaMemcpyHostToDevice, h->stream));
  if (E) {
    PCK(cudaMemcpyAsync(h->d_inc.p, inc.data(), 2 * E * 4, cudaMemcpyHostToDevice, h->stream));
    PCK(cudaMemcpyAsync(h->d_ea.p, h->ea.data(), E * 4, cudaMemcpyHostToDevice, h->stream)); PCK(cudaMemcpyAsync(h->d_eb.p, h->eb.data(), E * 4, cudaMemcpyHostToDevice, h->stream));
    PCK(cudaMemcpyAsync(h->d_rel.p, h->rel.data(), E * 64, cudaMemcpyHostToDevice, h->stream)); PCK(cudaMemcpyAsync(h->d_sinfo.p, h->sinfo.data(), E * 288, cudaMemcpyHostToDevice, h->stream));
  }
  PCK(cudaStreamSynchronize(h->stream));   // ptr / inc go out of scope
  h->uploaded = true;
  return 0;
}

static PgoDev pgo_view(d2pgo_handle *h, int cur) {
  PgoDev d; memset(&d, 0, sizeof d);
  d.n_pose = (int)h->ids.size(); d.n_edge = (int)h->ea.size(); d.x = h->d_x[cur].p; d.fixed = h->d_fixed.p; d.ea = h->d_ea.p; d.eb = h->d_eb.p;
  d.rel = h->d_rel.p; d.sinfo = h->d_sinfo.p; d.lin = h->d_lin[cur].p; d.Minv = h->d_Minv.p; d.dx = h->d_dx.p; d.r = h->d_r.p; d.z = h->d_z.p; d.p = h->d_p.p;
  d.Ap = h->d_Ap.p; d.s = h->d_s.p; d.damp = h->d_damp.p; d.t = h->d_t.p; d.inc_ptr = h->d_inc_ptr.p; d.inc = h->d_inc.p;
  d.nbp = (d.n_pose + 127) / 128; d.rz_part = h->d_part.p; d.rr_part = h->d_part.p + 2 * (size_t)d.nbp; d.pAp_part = h->d_part.p + 4 * (size_t)d.nbp;
  return d;
}

// cost (+ lin records, gradient, block diagonal of buffer `b`) at pose buffer `b`; all-reduced across the ranks
static int pgo_linearize(d2pgo_handle *h, int b, int want_jac, double *cost) {
  PgoDev d = pgo_view(h, b);
  const size_t N = h->ids.size();
  const int nbe = (d.n_edge + 127) / 128;
  if (d.n_edge > 0) k_pgo_lin<<<nbe, 128, 0, h->stream>>>(d, h->d_cost_part.p, want_jac);
  k_pgo_sum<<<1, 128, 0, h->stream>>>(h->d_cost_part.p, nbe, h->d_cost.p);
  if (want_jac) k_pgo_gD<<<d.nbp, 128, 0, h->stream>>>(d, h->d_g[b].p, h->d_D[b].p);
  if (h->comm) {
    if (nccl_allreduce_f64(h->comm, h->d_cost.p, 1, h->stream)) { h->err = "ncclAllReduce(cost) failed"; return 40; }
    if (want_jac && (nccl_allreduce_f64(h->comm, h->d_g[b].p, N * 6, h->stream) || nccl_allreduce_f64(h->comm, h->d_D[b].p, N * 36, h->stream))) { h->err = "ncclAllReduce(g, D) failed"; return 40; }
  }
  PCK(cudaMemcpyAsync(cost, h->d_cost.p, 8, cudaMemcpyDeviceToHost, h->stream));
  PCK(cudaStreamSynchronize(h->stream));
  return 0;
}

constexpr int kCgChunk = 16;   // CG iterations between two looks at the convergence flag (one graph launch on a single rank)

static int pgo_cg_chunk(d2pgo_handle *h, const PgoDev &d, double tol2) {
  const int ge = (d.n_edge + 127) / 128;
  for (int k = 0; k < kCgChunk; k++) {
    const int par = k & 1;
    if (d.n_edge > 0) k_pgo_cg_edge<<<ge, 128, 0, h->stream>>>(d, par, tol2);
    if (h->comm) {
      k_pgo_cg_pose<true><<<d.nbp, 128, 0, h->stream>>>(d, par, tol2);
      // every rank holds the same r, z, p (the all-reduced products are bitwise identical), so all of them reach the same
      // `done` decision at the same iteration and the collective below is always matched
      if (nccl_allreduce_f64(h->comm, d.Ap, (size_t)d.n_pose * 6, h->stream)) { h->err = "ncclAllReduce(Ap) failed"; return 40; }
      k_pgo_cg_pAp<<<d.nbp, 128, 0, h->stream>>>(d, par, tol2);
    } else k_pgo_cg_pose<false><<<d.nbp, 128, 0, h->stream>>>(d, par, tol2);
    k_pgo_cg_step<<<d.nbp, 128, 0, h->stream>>>(d, par, tol2);
  }
  return 0;
}

int d2pgo_solve(d2pgo_handle *h, d2pgo_report *rep) {
  if (!h || h->ids.empty()) return 1;
  cudaSetDevice(h->cfg.device);
  int rc;
  if (!h->uploaded && (rc = pgo_upload(h))) return rc;
  const int N = (int)h->ids.size();
  const int gp = (N + 127) / 128;
  d2pgo_report R; memset(&R, 0, sizeof R);
  PCK(cudaEventRecord(h->ev0, h->stream));
  int cur = 0;                    // buffer (poses, lin records, g, D) of the accepted point
  double cost = 0, lambda = h->cfg.lambda0;
  if ((rc = pgo_linearize(h, cur, 1, &cost))) return rc;
  R.initial_cost = cost;
  const double tol2 = h->cfg.pcg_tolerance * h->cfg.pcg_tolerance;
  if (tol2 != h->graph_tol2) { pgo_drop_graphs(h); h->graph_tol2 = tol2; }
  for (int it = 0; it < h->cfg.max_iterations; it++) {
    PgoDev d = pgo_view(h, cur);
    // (J^T J + lambda diag D) dx = -g by block-Jacobi preconditioned CG; J^T J is never formed
    k_pgo_precond<<<gp, 128, 0, h->stream>>>(d, h->d_D[cur].p, lambda);
    k_pgo_cg_init<<<gp, 128, 0, h->stream>>>(d, h->d_g[cur].p);
    k_pgo_cg_init2<<<1, 128, 0, h->stream>>>(d);
    PgoScalars s; memset(&s, 0, sizeof s);
    for (int k = 0; k < h->cfg.pcg_max_iterations; k += kCgChunk) {
      if (!h->comm) {
        if (!h->cg_graph[cur]) {
          cudaGraph_t g;
          PCK(cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal));
          pgo_cg_chunk(h, d, tol2);
          PCK(cudaStreamEndCapture(h->stream, &g));
          PCK(cudaGraphInstantiate(&h->cg_graph[cur], g, 0));
          cudaGraphDestroy(g);
        }
        PCK(cudaGraphLaunch(h->cg_graph[cur], h->stream));
      } else if ((rc = pgo_cg_chunk(h, d, tol2))) return rc;
      PCK(cudaMemcpyAsync(&s, h->d_s.p, sizeof s, cudaMemcpyDeviceToHost, h->stream));
      PCK(cudaStreamSynchronize(h->stream));
      if (s.done) break;   // identical on every rank (see pgo_cg_chunk)
    }
    k_pgo_retract<<<gp, 128, 0, h->stream>>>(d, h->d_x[1 - cur].p);
    double cand = 0;
    if ((rc = pgo_linearize(h, 1 - cur, 1, &cand))) return rc;
    R.pcg_iterations += s.iters; R.iterations++;
    if (cand < cost && isfinite(cand)) {
      const double rel_dec = (cost - cand) / (cost > 0 ? cost : 1.0);
      cur = 1 - cur; cost = cand; R.accepted++;
      lambda = lambda > 0 ? fmax(lambda / 3.0, 1e-12) : 0.0;
      if (rel_dec < h->cfg.function_tolerance) { R.converged = 1; break; }
    } else {
      lambda = lambda > 0 ? lambda * 4.0 : 1e-4;
      if (lambda > 1e8) break;
    }
  }
  PCK(cudaEventRecord(h->ev1, h->stream));
  PCK(cudaMemcpyAsync(h->poses.data(), h->d_x[cur].p, (size_t)N * 64, cudaMemcpyDeviceToHost, h->stream));
  PCK(cudaStreamSynchronize(h->stream));
  if (cur != 0) PCK(cudaMemcpy(h->d_x[0].p, h->d_x[1].p, (size_t)N * 64, cudaMemcpyDeviceToDevice));   // a following solve starts from buffer 0
  float ms = 0; cudaEventElapsedTime(&ms, h->ev0, h->ev1);
  R.final_cost = cost; R.device_ms = ms;
  if (rep) *rep = R;
  return 0;
}

int d2pgo_get_poses(d2pgo_handle *h, int32_t n, const int64_t *ids, double *out) {
  if (!h) return 1;
  for (int i = 0; i < n; i++) {
    auto it = h->index.find(ids[i]);
    if (it == h->index.end()) { h->err = "get_poses: unknown id"; return 2; }
    memcpy(out + (size_t)i * 7, &h->poses[(size_t)it->second * 8], 56);
  }
  return 0;
}

int d2pgo_debug_edges(d2pgo_handle *h, double *out, int64_t out_doubles) {
  if (!h) return 1;
  cudaSetDevice(h->cfg.device);
  int rc;
  if (!h->uploaded && (rc = pgo_upload(h))) return rc;
  const size_t E = h->ea.size();
  if ((size_t)out_doubles < E * 78) { h->err = "debug_edges: buffer too small"; return 2; }
  double cost;
  if ((rc = pgo_linearize(h, 0, 1, &cost))) return rc;
  std::vector<double> tmp(E * 78);
  PCK(cudaMemcpy(tmp.data(), h->d_lin[0].p, E * 78 * 8, cudaMemcpyDeviceToHost));
  for (size_t e = 0; e < E; e++) for (int k = 0; k < 78; k++) out[e * 78 + k] = tmp[(size_t)k * E + e];   // device layout is field-major
  return 0;
}

}  // extern "C"
