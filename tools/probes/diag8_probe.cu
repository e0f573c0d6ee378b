// Micro-benchmark of the 8x8 diagonal-block factorisation used by k_chol_smem (copy of chol_diag8, see d2ba_kernels.cu).
#include <cstdio>
#include <cuda_runtime.h>
#define D2BA_DEV __device__ __forceinline__
constexpr int kCsNB = 8;
D2BA_DEV double fast_rsqrt(double d) {
  double y = (double)rsqrtf((float)d);
  const double hd = 0.5 * d;
  y = y * (1.5 - hd * y * y);
  y = y * (1.5 - hd * y * y);
  return y;
}
// 1/d to full double precision: fp32 MUFU seed + two Newton steps; d must be a normal positive number in float range
D2BA_DEV double fast_rcp(double d) {
  float r32;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r32) : "f"((float)d));
  double r = (double)r32;
  double e = fma(-d, r, 1.0); r = fma(r, e, r);
  e = fma(-d, r, 1.0); r = fma(r, e, r);
  return r;
}
// 8x8 diagonal block at (k0, k0): lanes 0..7 of one warp hold one row each in registers.  Writes L_d back, 1/L_cc
// into invd and the row-scaled block M[c][k] = L[c][k] / L[c][c] (k < c) into Ms.  Returns true on a bad pivot.
D2BA_DEV bool chol_diag8(double *A, int ld, double *invd, double *Ms, int msld, int k0, int nb, int lane) {
  double row[kCsNB];
  const int r = k0 + lane;
#pragma unroll
  for (int c = 0; c < kCsNB; c++) row[c] = (lane < nb && c <= lane) ? A[(size_t)r * ld + k0 + c] : (c == lane ? 1.0 : 0.0);
  bool bad = false;
  double dmine = 1.0;   // pivot of this lane's own column
#pragma unroll
  for (int c = 0; c < kCsNB; c++) {
    const double dcc = __shfl_sync(0xffffffffu, row[c], c);
    const bool live = c < nb;
    const bool pos = dcc > 1e-30 && dcc < 1e30;
    if (live && !pos) bad = true;
    if (lane == c) dmine = dcc;
    const double uc = row[c];   // unscaled entry of this lane in column c
    double pr[kCsNB];
#pragma unroll
    for (int c2 = c + 1; c2 < kCsNB; c2++) pr[c2] = uc * __shfl_sync(0xffffffffu, uc, c2);   // independent of the reciprocal
    const double rc = (live && pos) ? fast_rcp(dcc) : 0.0;
#pragma unroll
    for (int c2 = c + 1; c2 < kCsNB; c2++)
      if (c2 <= lane) row[c2] = fma(-pr[c2], rc, row[c2]);
  }
  // scale: L[r][c] = U[r][c] / sqrt(d_c)
  const double smine = (lane < nb && dmine > 1e-30 && dmine < 1e30) ? fast_rsqrt(dmine) : 1.0;
  double sc[kCsNB];
#pragma unroll
  for (int c = 0; c < kCsNB; c++) sc[c] = __shfl_sync(0xffffffffu, smine, c);
  if (lane < nb) {
    invd[k0 + lane] = smine;
#pragma unroll
    for (int c = 0; c < kCsNB; c++) {
      const double l = row[c] * sc[c];
      if (c <= lane) A[(size_t)r * ld + k0 + c] = l;
      Ms[lane * msld + c] = (c < lane) ? l * smine : 0.0;
    }
  } else if (lane < kCsNB) {
#pragma unroll
    for (int c = 0; c < kCsNB; c++) Ms[lane * msld + c] = 0.0;
  }
  return bad;
}


__global__ void k_diag(double *out, int iters, long long *cyc) {
  __shared__ double A[8 * 8], invd[8], Ms[64], A0[64];
  const int lane = threadIdx.x;
  if (lane < 8) for (int c = 0; c < 8; c++) A0[lane * 8 + c] = (c == lane ? 10.0 + lane : 1.0 / (1 + lane + c));
  __syncwarp();
  long long t0 = 0, tot = 0; bool bad = false;
  for (int it = 0; it < iters; it++) {
    for (int e = lane; e < 64; e += 32) A[e] = A0[e];
    __syncwarp();
    t0 = clock64();
    bad |= chol_diag8(A, 8, invd, Ms, 8, 0, 8, lane);
    __syncwarp();
    tot += clock64() - t0;
  }
  if (lane == 0) { *cyc = tot; out[0] = A[63] + invd[7] + Ms[60] + (bad ? 1 : 0); }
}
int main() {
  double *out; long long *cyc, h; cudaMalloc(&out, 64); cudaMalloc(&cyc, 8);
  k_diag<<<1, 32>>>(out, 1000, cyc); cudaDeviceSynchronize();
  k_diag<<<1, 32>>>(out, 1000, cyc); cudaDeviceSynchronize();
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  double o; cudaMemcpy(&o, out, 8, cudaMemcpyDeviceToHost);
  printf("chol_diag8: %.1f cycles per call (check %.6f) %s\n", h / 1000.0, o, cudaGetErrorString(cudaGetLastError()));
  return 0;
}
