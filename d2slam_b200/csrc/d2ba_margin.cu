// d2ba_margin.cu -- prior construction on the device.
//
//   toJacRes  (d2vins/src/factors/prior_factor.cpp:132-177): A <- (A + A^T)/2, A = V diag(s) V^T,
//             s <= 1e-8 -> 0, J = sqrt(S) V^T, e0 = sqrt(S^-1) V^T b.
// The symmetric eigen-decomposition is a one-CTA parallel-ordered (round-robin) two-sided Jacobi in
// fp64: each step applies m/2 disjoint rotations, columns first and then rows.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/d2ba.h"

namespace d2ba {

constexpr int kEigThreads = 1024;

// A (mp x mp, mp even, padded with zero rows/cols and a distinct negative diagonal) and V in global memory.
__global__ void __launch_bounds__(kEigThreads) k_sym_eig(double *A, double *V, int mp, int max_sweeps, double *cs /*2*mp*/, int *pq /*2*mp*/) {
  const int tid = threadIdx.x, nt = blockDim.x;
  const int half = mp / 2;
  __shared__ double s_off, s_diag;
  for (int e = tid; e < mp * mp; e += nt) V[e] = (e / mp == e % mp) ? 1.0 : 0.0;
  __syncthreads();
  for (int sweep = 0; sweep < max_sweeps; sweep++) {
    // convergence: off-diagonal Frobenius mass
    if (tid == 0) { s_off = 0; s_diag = 0; }
    __syncthreads();
    double off = 0, dg = 0;
    for (int e = tid; e < mp * mp; e += nt) { int i = e / mp, j = e % mp; double v = A[e]; if (i == j) dg += v * v; else off += v * v; }
    for (int o = 16; o > 0; o >>= 1) { off += __shfl_xor_sync(0xffffffffu, off, o); dg += __shfl_xor_sync(0xffffffffu, dg, o); }
    if ((tid & 31) == 0) { atomicAdd(&s_off, off); atomicAdd(&s_diag, dg); }
    __syncthreads();
    if (s_off <= 1e-30 * (s_diag + s_off) || s_off == 0.0) break;
    for (int step = 0; step < mp - 1; step++) {
      // round-robin pairing: player mp-1 fixed, others rotate
      for (int k = tid; k < half; k += nt) {
        int a = (k == 0) ? mp - 1 : (step + k) % (mp - 1);
        int b = (step + mp - 1 - k) % (mp - 1);
        int p = a < b ? a : b, q = a < b ? b : a;
        double apq = A[p * mp + q], c = 1.0, s = 0.0;
        if (apq != 0.0) {
          double tau = (A[q * mp + q] - A[p * mp + p]) / (2.0 * apq);
          double t = (tau >= 0 ? 1.0 : -1.0) / (fabs(tau) + sqrt(1.0 + tau * tau));
          c = 1.0 / sqrt(1.0 + t * t); s = t * c;
        }
        cs[2 * k] = c; cs[2 * k + 1] = s; pq[2 * k] = p; pq[2 * k + 1] = q;
      }
      __syncthreads();
      // columns: A <- A J, V <- V J
      for (int e = tid; e < half * mp; e += nt) {
        int k = e / mp, r = e % mp;
        double c = cs[2 * k], s = cs[2 * k + 1];
        int p = pq[2 * k], q = pq[2 * k + 1];
        double ap = A[r * mp + p], aq = A[r * mp + q];
        A[r * mp + p] = c * ap - s * aq; A[r * mp + q] = s * ap + c * aq;
        double vp = V[r * mp + p], vq = V[r * mp + q];
        V[r * mp + p] = c * vp - s * vq; V[r * mp + q] = s * vp + c * vq;
      }
      __syncthreads();
      // rows: A <- J^T A
      for (int e = tid; e < half * mp; e += nt) {
        int k = e / mp, r = e % mp;
        double c = cs[2 * k], s = cs[2 * k + 1];
        int p = pq[2 * k], q = pq[2 * k + 1];
        double ap = A[p * mp + r], aq = A[q * mp + r];
        A[p * mp + r] = c * ap - s * aq; A[q * mp + r] = s * ap + c * aq;
      }
      __syncthreads();
    }
  }
}

// J[i][k] = sqrt(s_i) V[k][i], e0[i] = sqrt(1/s_i) sum_k V[k][i] b[k]   (eigenvalue i <= eps -> 0)
__global__ void k_to_jac_res(const double *A, const double *V, int mp, int m, const double *b, double *J, double *e0) {
  int i = blockIdx.x;
  if (i >= m) return;
  // eigenpairs of the padding block have (strongly) negative eigenvalues and are clamped away; to keep J m x m we
  // emit the m rows with the largest eigenvalues via a rank computed on the fly
  __shared__ int src;
  __shared__ double red[32];
  if (threadIdx.x == 0) {
    // i-th largest eigenvalue index (m <= few hundred: O(m^2) total is fine)
    int found = -1;
    for (int cand = 0; cand < mp && found < 0; cand++) {
      int rank = 0;
      double ev = A[cand * mp + cand];
      for (int o = 0; o < mp; o++) { double eo = A[o * mp + o]; if (eo > ev || (eo == ev && o < cand)) rank++; }
      if (rank == i) found = cand;
    }
    src = found;
  }
  __syncthreads();
  const int c = src;
  const double ev = A[c * mp + c], eps = 1e-8;
  const double S = ev > eps ? ev : 0.0, Si = ev > eps ? 1.0 / ev : 0.0;
  const double ss = sqrt(S), si = sqrt(Si);
  double dot = 0;
  for (int k = threadIdx.x; k < m; k += blockDim.x) { double v = V[k * mp + c]; J[(size_t)i * m + k] = ss * v; dot += v * b[k]; }
  for (int o = 16; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = dot;
  __syncthreads();
  if (threadIdx.x == 0) { double t = 0; for (int q = 0; q < (int)(blockDim.x >> 5); q++) t += red[q]; e0[i] = si * t; }
}

}  // namespace d2ba

namespace d2ba {
// Batched toJacRes: one CTA per window whose prior was given in information form.  In place:
// Aio (m x m) holds A on entry and J on exit; bio (m) holds b on entry and e0 on exit; V is scratch (m x m).
__global__ void __launch_bounds__(kEigThreads) k_prior_from_info(const int *m_of, const long long *offJ, const long long *offv,
                                                                const int *is_info, double *Aall, double *Vall, double *ball) {
  const int wi = blockIdx.x;
  const int m = m_of[wi];
  if (m <= 0 || !is_info[wi]) return;
  double *A = Aall + offJ[wi], *V = Vall + offJ[wi], *b = ball + offv[wi];
  const int tid = threadIdx.x, nt = blockDim.x;
  extern __shared__ double smx[];
  double *cs = smx;                                  // 2 * half
  int *pq = reinterpret_cast<int *>(smx + m + 2);    // 2 * half ints
  double *ev = smx + 2 * (m + 2);                    // m eigenvalues
  double *bb = ev + m;                               // m
  __shared__ double s_off, s_diag;
  const int mp = (m + 1) / 2 * 2, half = mp / 2;
  // symmetrise, V = I
  for (int e = tid; e < m * m; e += nt) { int i = e / m, j = e % m; if (j < i) { double v = 0.5 * (A[i * m + j] + A[j * m + i]); A[i * m + j] = v; A[j * m + i] = v; } V[e] = (i == j) ? 1.0 : 0.0; }
  for (int i = tid; i < m; i += nt) bb[i] = b[i];
  __syncthreads();
  for (int sweep = 0; sweep < 60; sweep++) {
    if (tid == 0) { s_off = 0; s_diag = 0; }
    __syncthreads();
    double off = 0, dg = 0;
    for (int e = tid; e < m * m; e += nt) { int i = e / m, j = e % m; double v = A[e]; if (i == j) dg += v * v; else off += v * v; }
    for (int o = 16; o > 0; o >>= 1) { off += __shfl_xor_sync(0xffffffffu, off, o); dg += __shfl_xor_sync(0xffffffffu, dg, o); }
    if ((tid & 31) == 0) { atomicAdd(&s_off, off); atomicAdd(&s_diag, dg); }
    __syncthreads();
    if (s_off <= 1e-30 * (s_diag + s_off) || s_off == 0.0) break;
    for (int step = 0; step < mp - 1; step++) {
      for (int k = tid; k < half; k += nt) {
        int a = (k == 0) ? mp - 1 : (step + k) % (mp - 1);
        int bq = (step + mp - 1 - k) % (mp - 1);
        int p = a < bq ? a : bq, q = a < bq ? bq : a;
        double c = 1.0, s = 0.0;
        if (q < m) {
          double apq = A[p * m + q];
          if (apq != 0.0) {
            double tau = (A[q * m + q] - A[p * m + p]) / (2.0 * apq);
            double t = (tau >= 0 ? 1.0 : -1.0) / (fabs(tau) + sqrt(1.0 + tau * tau));
            c = 1.0 / sqrt(1.0 + t * t); s = t * c;
          }
        } else { q = -1; }
        cs[2 * k] = c; cs[2 * k + 1] = s; pq[2 * k] = p; pq[2 * k + 1] = q;
      }
      __syncthreads();
      for (int e = tid; e < half * m; e += nt) {
        int k = e / m, r = e % m;
        int p = pq[2 * k], q = pq[2 * k + 1];
        if (q < 0) continue;
        double c = cs[2 * k], s = cs[2 * k + 1];
        double ap = A[r * m + p], aq = A[r * m + q];
        A[r * m + p] = c * ap - s * aq; A[r * m + q] = s * ap + c * aq;
        double vp = V[r * m + p], vq = V[r * m + q];
        V[r * m + p] = c * vp - s * vq; V[r * m + q] = s * vp + c * vq;
      }
      __syncthreads();
      for (int e = tid; e < half * m; e += nt) {
        int k = e / m, r = e % m;
        int p = pq[2 * k], q = pq[2 * k + 1];
        if (q < 0) continue;
        double c = cs[2 * k], s = cs[2 * k + 1];
        double ap = A[p * m + r], aq = A[q * m + r];
        A[p * m + r] = c * ap - s * aq; A[q * m + r] = s * ap + c * aq;
      }
      __syncthreads();
    }
  }
  for (int i = tid; i < m; i += nt) ev[i] = A[i * m + i];
  __syncthreads();
  // J[i][k] = sqrt(s_i) V[k][i];  e0[i] = sqrt(1/s_i) sum_k V[k][i] b[k]   (prior_factor.cpp:139-150)
  const double eps = 1e-8;
  for (int e = tid; e < m * m; e += nt) { int i = e / m, k = e % m; double s = ev[i] > eps ? sqrt(ev[i]) : 0.0; A[e] = s * V[k * m + i]; }
  for (int i = tid; i < m; i += nt) {
    double dot = 0;
    for (int k = 0; k < m; k++) dot += V[k * m + i] * bb[k];
    b[i] = (ev[i] > eps ? sqrt(1.0 / ev[i]) : 0.0) * dot;
  }
}

void launch_prior_from_info(int n_win, int max_m, const int *m_of, const long long *offJ, const long long *offv, const int *is_info,
                            double *A, double *V, double *b, cudaStream_t s) {
  size_t sm = (size_t)(4 * (max_m + 2) + 8) * 8;
  k_prior_from_info<<<n_win, kEigThreads, sm, s>>>(m_of, offJ, offv, is_info, A, V, b);
}
}  // namespace d2ba

extern "C" int d2ba_prior_info_to_jac(d2ba_handle *h, int m, const double *A, const double *b, double *J, double *e0) {
  (void)h;
  if (m <= 0) return 0;
  const int mp = (m + 1) / 2 * 2;
  std::vector<double> Ap((size_t)mp * mp, 0.0);
  for (int i = 0; i < m; i++) for (int j = 0; j < m; j++) Ap[(size_t)i * mp + j] = 0.5 * (A[(size_t)i * m + j] + A[(size_t)j * m + i]);
  if (mp > m) Ap[(size_t)m * mp + m] = -1.0;   // padding eigenvalue, clamped away
  double *dA = nullptr, *dV = nullptr, *dcs = nullptr, *db = nullptr, *dJ = nullptr, *de0 = nullptr; int *dpq = nullptr;
  cudaError_t e = cudaSuccess;
  auto ok = [&](cudaError_t x) { if (e == cudaSuccess) e = x; };
  ok(cudaMalloc(&dA, Ap.size() * 8)); ok(cudaMalloc(&dV, Ap.size() * 8)); ok(cudaMalloc(&dcs, (size_t)mp * 2 * 8)); ok(cudaMalloc(&dpq, (size_t)mp * 2 * 4));
  ok(cudaMalloc(&db, (size_t)m * 8)); ok(cudaMalloc(&dJ, (size_t)m * m * 8)); ok(cudaMalloc(&de0, (size_t)m * 8));
  if (e == cudaSuccess) {
    ok(cudaMemcpy(dA, Ap.data(), Ap.size() * 8, cudaMemcpyHostToDevice)); ok(cudaMemcpy(db, b, (size_t)m * 8, cudaMemcpyHostToDevice));
    d2ba::k_sym_eig<<<1, d2ba::kEigThreads>>>(dA, dV, mp, 60, dcs, dpq);
    d2ba::k_to_jac_res<<<m, 128>>>(dA, dV, mp, m, db, dJ, de0);
    ok(cudaGetLastError());
    ok(cudaMemcpy(J, dJ, (size_t)m * m * 8, cudaMemcpyDeviceToHost)); ok(cudaMemcpy(e0, de0, (size_t)m * 8, cudaMemcpyDeviceToHost));
  }
  cudaFree(dA); cudaFree(dV); cudaFree(dcs); cudaFree(dpq); cudaFree(db); cudaFree(dJ); cudaFree(de0);
  return e == cudaSuccess ? 0 : 200 + (int)e;
}

extern "C" int d2ba_marginalize(d2ba_handle *h, int32_t window, int32_t n_remove, const int64_t *remove_frame_ids,
                                int32_t *m_out, int32_t max_m, double *A_out, double *b_out, int32_t *nblk_out,
                                int32_t max_blk, d2ba_blockref *refs_out) {
  (void)h; (void)window; (void)n_remove; (void)remove_frame_ids; (void)m_out; (void)max_m; (void)A_out; (void)b_out;
  (void)nblk_out; (void)max_blk; (void)refs_out;
  return 100;   // built after rows (a)-(e)
}
