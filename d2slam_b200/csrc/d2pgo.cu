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

struct PgoScalars {      // device-resident CG / LM scalars
  double rz, rz_new, pAp, rr, bb, cost, rr_last;
  int done, iters;
};

struct PgoDev {
  int n_pose, n_edge;
  const double *x;        // [N][8] poses the edges are linearised at
  const unsigned char *fixed;
  const int *ea, *eb;     // [E] pose indices
  const double *rel;      // [E][8]
  const double *sinfo;    // [E][36] sqrt information, row-major
  double *lin;            // [E][78]: r(6), J0 (6x6 row-major), J1 (6x6)
  double *g, *D;          // [6N], [N][36]
  double *Minv;           // [N][36]
  double *dx, *r, *z, *p, *Ap;   // [6N]
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

// linearise every local edge at d.x: lin records, cost, gradient g = J^T r and the block diagonal D = sum J^T J
__global__ void __launch_bounds__(128) k_pgo_lin(PgoDev d, double *g, double *D, double *cost_out, int want_jac) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  __shared__ double red[40];
  double c = 0.0;
  if (e < d.n_edge) {
    const int a = d.ea[e], b = d.eb[e];
    double r[6], J0[36], J1[36];
    pgo_edge_eval(d.x + (size_t)a * 8, d.x + (size_t)b * 8, d.rel + (size_t)e * 8, d.sinfo + (size_t)e * 36, r, want_jac ? J0 : nullptr, J1);
    for (int k = 0; k < 6; k++) c += 0.5 * r[k] * r[k];
    if (want_jac) {
      double *o = d.lin + (size_t)e * 78;
      for (int k = 0; k < 6; k++) o[k] = r[k];
      for (int k = 0; k < 36; k++) { o[6 + k] = J0[k]; o[42 + k] = J1[k]; }
      const bool fa = d.fixed[a], fb = d.fixed[b];
      for (int i = 0; i < 6; i++) {
        double ga = 0, gb = 0;
        for (int k = 0; k < 6; k++) { ga += J0[k * 6 + i] * r[k]; gb += J1[k * 6 + i] * r[k]; }
        if (!fa) atomicAdd(&g[(size_t)a * 6 + i], ga);
        if (!fb) atomicAdd(&g[(size_t)b * 6 + i], gb);
        for (int j = 0; j < 6; j++) {
          double ha = 0, hb = 0;
          for (int k = 0; k < 6; k++) { ha += J0[k * 6 + i] * J0[k * 6 + j]; hb += J1[k * 6 + i] * J1[k * 6 + j]; }
          if (!fa && ha != 0.0) atomicAdd(&D[(size_t)a * 36 + i * 6 + j], ha);
          if (!fb && hb != 0.0) atomicAdd(&D[(size_t)b * 36 + i * 6 + j], hb);
        }
      }
    }
  }
  c = block_sum(c, red);
  if (threadIdx.x == 0 && c != 0.0) atomicAdd(cost_out, c);
}

// block-Jacobi preconditioner: Minv = (D + lambda diag(D))^-1 per free pose (6x6 Cholesky, one thread per pose)
__global__ void k_pgo_precond(PgoDev d, const double *D, double lambda) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= d.n_pose) return;
  double *M = d.Minv + (size_t)i * 36;
  if (d.fixed[i]) { for (int k = 0; k < 36; k++) M[k] = 0.0; return; }
  double A[36], Li[36];
  for (int k = 0; k < 36; k++) A[k] = D[(size_t)i * 36 + k];
  for (int k = 0; k < 6; k++) A[k * 7] += lambda * A[k * 7] + 1e-12;
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

// CG start: dx = 0, r = -g, z = Minv r, p = z, scalars
__global__ void k_pgo_cg_init(PgoDev d, const double *g) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  __shared__ double red[40];
  double rz = 0, rr = 0;
  if (i < d.n_pose) {
    double r[6], z[6];
    for (int k = 0; k < 6; k++) r[k] = d.fixed[i] ? 0.0 : -g[(size_t)i * 6 + k];
    const double *M = d.Minv + (size_t)i * 36;
    for (int k = 0; k < 6; k++) { double t = 0; for (int q = 0; q < 6; q++) t += M[k * 6 + q] * r[q]; z[k] = t; }
    for (int k = 0; k < 6; k++) {
      d.dx[(size_t)i * 6 + k] = 0.0; d.r[(size_t)i * 6 + k] = r[k]; d.z[(size_t)i * 6 + k] = z[k]; d.p[(size_t)i * 6 + k] = z[k]; d.Ap[(size_t)i * 6 + k] = 0.0;
      rz += r[k] * z[k]; rr += r[k] * r[k];
    }
  }
  rz = block_sum(rz, red); rr = block_sum(rr, red);
  if (threadIdx.x == 0) { atomicAdd(&d.s->rz, rz); atomicAdd(&d.s->bb, rr); }
}

// Ap += J^T (J p) over the local edges (matrix-free)
__global__ void __launch_bounds__(128) k_pgo_matvec(PgoDev d) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= d.n_edge || d.s->done) return;
  const int a = d.ea[e], b = d.eb[e];
  const double *J0 = d.lin + (size_t)e * 78 + 6, *J1 = J0 + 36;
  double pa[6], pb[6], t[6];
  for (int k = 0; k < 6; k++) { pa[k] = d.p[(size_t)a * 6 + k]; pb[k] = d.p[(size_t)b * 6 + k]; }
  for (int k = 0; k < 6; k++) { double s = 0; for (int q = 0; q < 6; q++) s += J0[k * 6 + q] * pa[q] + J1[k * 6 + q] * pb[q]; t[k] = s; }
  const bool fa = d.fixed[a], fb = d.fixed[b];
  for (int q = 0; q < 6; q++) {
    double ya = 0, yb = 0;
    for (int k = 0; k < 6; k++) { ya += J0[k * 6 + q] * t[k]; yb += J1[k * 6 + q] * t[k]; }
    if (!fa) atomicAdd(&d.Ap[(size_t)a * 6 + q], ya);
    if (!fb) atomicAdd(&d.Ap[(size_t)b * 6 + q], yb);
  }
}

// damping + p.Ap
__global__ void k_pgo_cg_damp(PgoDev d, const double *D, double lambda) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  __shared__ double red[40];
  double s = 0;
  if (i < d.n_pose && !d.s->done && !d.fixed[i])
    for (int k = 0; k < 6; k++) {
      const double pk = d.p[(size_t)i * 6 + k];
      const double ap = d.Ap[(size_t)i * 6 + k] + (lambda * D[(size_t)i * 36 + k * 7] + 1e-12) * pk;
      d.Ap[(size_t)i * 6 + k] = ap; s += pk * ap;
    }
  s = block_sum(s, red);
  if (threadIdx.x == 0 && s != 0.0) atomicAdd(&d.s->pAp, s);
}

// dx += alpha p; r -= alpha Ap; z = Minv r; r.z, r.r; Ap = 0 for the next product
__global__ void k_pgo_cg_update(PgoDev d) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  __shared__ double red[40];
  double rz = 0, rr = 0;
  if (i < d.n_pose && !d.s->done && !d.fixed[i]) {
    const double alpha = d.s->rz / d.s->pAp;
    double r[6];
    for (int k = 0; k < 6; k++) {
      d.dx[(size_t)i * 6 + k] += alpha * d.p[(size_t)i * 6 + k];
      r[k] = d.r[(size_t)i * 6 + k] - alpha * d.Ap[(size_t)i * 6 + k];
      d.r[(size_t)i * 6 + k] = r[k]; d.Ap[(size_t)i * 6 + k] = 0.0;
    }
    const double *M = d.Minv + (size_t)i * 36;
    for (int k = 0; k < 6; k++) { double t = 0; for (int q = 0; q < 6; q++) t += M[k * 6 + q] * r[q]; d.z[(size_t)i * 6 + k] = t; rz += r[k] * t; rr += r[k] * r[k]; }
  }
  rz = block_sum(rz, red); rr = block_sum(rr, red);
  if (threadIdx.x == 0) { if (rz != 0.0) atomicAdd(&d.s->rz_new, rz); if (rr != 0.0) atomicAdd(&d.s->rr, rr); }
}

// p = z + beta p
__global__ void k_pgo_cg_dir(PgoDev d) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= d.n_pose * 6 || d.s->done) return;
  const double beta = d.s->rz_new / d.s->rz;
  d.p[i] = d.z[i] + beta * d.p[i];
}

// rotate the scalars, convergence test
__global__ void k_pgo_cg_next(PgoDev d, double tol2) {
  PgoScalars *s = d.s;
  if (s->done) return;
  s->iters++;
  // tol2 < 0 (multi-rank): the ranks decide together on the host from all-reduced values, never from this rank's own flag
  if ((tol2 >= 0.0 && !(s->rr > tol2 * s->bb)) || !(s->rz_new > 0.0) || !isfinite(s->rz_new)) s->done = 1;
  s->rr_last = s->rr;
  s->rz = s->rz_new; s->rz_new = 0.0; s->pAp = 0.0; s->rr = 0.0;
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
  Buf<double> d_x[2], d_rel, d_sinfo, d_lin, d_g[2], d_D[2], d_Minv, d_dx, d_r, d_z, d_p, d_Ap, d_cost;
  Buf<unsigned char> d_fixed; Buf<int> d_ea, d_eb; Buf<PgoScalars> d_s;
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
  for (int b = 0; b < 2; b++) { h->d_x[b].release(); h->d_g[b].release(); h->d_D[b].release(); }
  h->d_rel.release(); h->d_sinfo.release(); h->d_lin.release(); h->d_Minv.release(); h->d_dx.release(); h->d_r.release(); h->d_z.release(); h->d_p.release();
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

static int pgo_upload(d2pgo_handle *h) {
  const size_t N = h->ids.size(), E = h->ea.size();
  for (int b = 0; b < 2; b++) { PCK(h->d_x[b].alloc(N * 8)); PCK(h->d_g[b].alloc(N * 6)); PCK(h->d_D[b].alloc(N * 36)); }
  PCK(h->d_rel.alloc(E * 8)); PCK(h->d_sinfo.alloc(E * 36)); PCK(h->d_lin.alloc(E * 78)); PCK(h->d_Minv.alloc(N * 36));
  PCK(h->d_dx.alloc(N * 6)); PCK(h->d_r.alloc(N * 6)); PCK(h->d_z.alloc(N * 6)); PCK(h->d_p.alloc(N * 6)); PCK(h->d_Ap.alloc(N * 6)); PCK(h->d_cost.alloc(2));
  PCK(h->d_fixed.alloc(N)); PCK(h->d_ea.alloc(E)); PCK(h->d_eb.alloc(E)); PCK(h->d_s.alloc(1));
  PCK(cudaMemcpyAsync(h->d_x[0].p, h->poses.data(), N * 64, cudaMemcpyHostToDevice, h->stream));
  PCK(cudaMemcpyAsync(h->d_fixed.p, h->fixed.data(), N, cudaMemcpyHostToDevice, h->stream));
  if (E) {
    PCK(cudaMemcpyAsync(h->d_ea.p, h->ea.data(), E * 4, cudaMemcpyHostToDevice, h->stream)); PCK(cudaMemcpyAsync(h->d_eb.p, h->eb.data(), E * 4, cudaMemcpyHostToDevice, h->stream));
    PCK(cudaMemcpyAsync(h->d_rel.p, h->rel.data(), E * 64, cudaMemcpyHostToDevice, h->stream)); PCK(cudaMemcpyAsync(h->d_sinfo.p, h->sinfo.data(), E * 288, cudaMemcpyHostToDevice, h->stream));
  }
  h->uploaded = true;
  return 0;
}

static PgoDev pgo_view(d2pgo_handle *h, int cur) {
  PgoDev d; memset(&d, 0, sizeof d);
  d.n_pose = (int)h->ids.size(); d.n_edge = (int)h->ea.size(); d.x = h->d_x[cur].p; d.fixed = h->d_fixed.p; d.ea = h->d_ea.p; d.eb = h->d_eb.p;
  d.rel = h->d_rel.p; d.sinfo = h->d_sinfo.p; d.lin = h->d_lin.p; d.Minv = h->d_Minv.p; d.dx = h->d_dx.p; d.r = h->d_r.p; d.z = h->d_z.p; d.p = h->d_p.p;
  d.Ap = h->d_Ap.p; d.s = h->d_s.p;
  return d;
}

// cost (+ gradient / block diagonal into buffer `gb`) at pose buffer `xb`; all-reduced across the ranks
static int pgo_linearize(d2pgo_handle *h, int xb, int gb, int want_jac, double *cost) {
  PgoDev d = pgo_view(h, xb);
  const size_t N = h->ids.size();
  PCK(cudaMemsetAsync(h->d_cost.p, 0, 16, h->stream));
  if (want_jac) { PCK(cudaMemsetAsync(h->d_g[gb].p, 0, N * 48, h->stream)); PCK(cudaMemsetAsync(h->d_D[gb].p, 0, N * 288, h->stream)); }
  if (d.n_edge > 0) k_pgo_lin<<<(d.n_edge + 127) / 128, 128, 0, h->stream>>>(d, h->d_g[gb].p, h->d_D[gb].p, h->d_cost.p, want_jac);
  if (h->comm) {
    if (nccl_allreduce_f64(h->comm, h->d_cost.p, 1, h->stream)) { h->err = "ncclAllReduce(cost) failed"; return 40; }
    if (want_jac && (nccl_allreduce_f64(h->comm, h->d_g[gb].p, N * 6, h->stream) || nccl_allreduce_f64(h->comm, h->d_D[gb].p, N * 36, h->stream))) { h->err = "ncclAllReduce(g, D) failed"; return 40; }
  }
  PCK(cudaMemcpyAsync(cost, h->d_cost.p, 8, cudaMemcpyDeviceToHost, h->stream));
  PCK(cudaStreamSynchronize(h->stream));
  return 0;
}

int d2pgo_solve(d2pgo_handle *h, d2pgo_report *rep) {
  if (!h || h->ids.empty()) return 1;
  cudaSetDevice(h->cfg.device);
  int rc;
  if (!h->uploaded && (rc = pgo_upload(h))) return rc;
  const int N = (int)h->ids.size(), E = (int)h->ea.size();
  const int gp = (N + 127) / 128, ge = (E + 127) / 128, gv = (N * 6 + 255) / 256;
  d2pgo_report R; memset(&R, 0, sizeof R);
  PCK(cudaEventRecord(h->ev0, h->stream));
  int cur = 0;                    // pose buffer / (g, D) buffer of the accepted point
  double cost = 0, lambda = h->cfg.lambda0;
  if ((rc = pgo_linearize(h, cur, cur, 1, &cost))) return rc;
  R.initial_cost = cost;
  for (int it = 0; it < h->cfg.max_iterations; it++) {
    PgoDev d = pgo_view(h, cur);
    const double *g = h->d_g[cur].p, *D = h->d_D[cur].p;
    // (J^T J + lambda diag D) dx = -g by block-Jacobi preconditioned CG; J^T J is never formed
    k_pgo_precond<<<gp, 128, 0, h->stream>>>(d, D, lambda);
    PCK(cudaMemsetAsync(h->d_s.p, 0, sizeof(PgoScalars), h->stream));
    k_pgo_cg_init<<<gp, 128, 0, h->stream>>>(d, g);
    const double tol2 = h->cfg.pcg_tolerance * h->cfg.pcg_tolerance;
    for (int k = 0; k < h->cfg.pcg_max_iterations; k++) {
      if (E > 0) k_pgo_matvec<<<ge, 128, 0, h->stream>>>(d);
      if (h->comm && nccl_allreduce_f64(h->comm, h->d_Ap.p, (size_t)N * 6, h->stream)) { h->err = "ncclAllReduce(Ap) failed"; return 40; }
      k_pgo_cg_damp<<<gp, 128, 0, h->stream>>>(d, D, lambda);
      k_pgo_cg_update<<<gp, 128, 0, h->stream>>>(d);
      k_pgo_cg_dir<<<gv, 256, 0, h->stream>>>(d);
      k_pgo_cg_next<<<1, 1, 0, h->stream>>>(d, h->comm ? -1.0 : tol2);
      if ((k & 15) == 15) {   // every 16 iterations: converged?
        PgoScalars s;
        if (h->comm) {
          // every rank must leave the loop at the same iteration (the all-reduce inside it is collective): decide from
          // all-reduced residual norms, which are bitwise identical on all ranks
          PCK(cudaMemcpyAsync(h->d_cost.p, &d.s->rr_last, 8, cudaMemcpyDeviceToDevice, h->stream));
          PCK(cudaMemcpyAsync(h->d_cost.p + 1, &d.s->bb, 8, cudaMemcpyDeviceToDevice, h->stream));
          if (nccl_allreduce_f64(h->comm, h->d_cost.p, 2, h->stream)) { h->err = "ncclAllReduce(residual) failed"; return 40; }
          double v[2];
          PCK(cudaMemcpyAsync(v, h->d_cost.p, 16, cudaMemcpyDeviceToHost, h->stream));
          PCK(cudaStreamSynchronize(h->stream));
          if (!(v[0] > tol2 * v[1])) break;
        } else {
          PCK(cudaMemcpyAsync(&s, h->d_s.p, sizeof s, cudaMemcpyDeviceToHost, h->stream));
          PCK(cudaStreamSynchronize(h->stream));
          if (s.done) break;
        }
      }
    }
    PgoScalars s;
    PCK(cudaMemcpyAsync(&s, h->d_s.p, sizeof s, cudaMemcpyDeviceToHost, h->stream));
    k_pgo_retract<<<gp, 128, 0, h->stream>>>(d, h->d_x[1 - cur].p);
    double cand = 0;
    if ((rc = pgo_linearize(h, 1 - cur, 1 - cur, 1, &cand))) return rc;
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
  if ((rc = pgo_linearize(h, 0, 0, 1, &cost))) return rc;
  PCK(cudaMemcpy(out, h->d_lin.p, E * 78 * 8, cudaMemcpyDeviceToHost));
  return 0;
}

}  // extern "C"
