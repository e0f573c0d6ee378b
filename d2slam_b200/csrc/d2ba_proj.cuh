// d2ba_proj.cuh -- reprojection factor arithmetic on the device.
//
// Computes the same residuals / Jacobians as the reference factors
//   ProjectionTwoFrameOneCamFactor::Evaluate        d2vins/src/factors/projectionTwoFrameOneCamFactor.cpp:48-177
//   ProjectionTwoFrameTwoCamFactor::Evaluate        d2vins/src/factors/projectionTwoFrameTwoCamFactor.cpp:46-187
//   ProjectionOneFrameTwoCamFactor::Evaluate        d2vins/src/factors/projectionOneFrameTwoCamFactor.cpp:46-160
//   ProjectionTwoFrameOneCamDepthFactor::Evaluate   d2vins/src/factors/projectionTwoFrameOneCamDepthFactor.cpp:47-186
// but organised for the GPU: everything that depends only on the parameter blocks of a residual
// group (rotation products J_w, J_m, J_c, the relative transform) is computed once per group by the
// warp and kept in shared memory; per observation only 2x3 row-vector products remain.
// Reference quirks reproduced on purpose: unit-sphere tangent residual, td Jacobian's
// mixed-norm matrix, single-camera extrinsic rotation block's Rj^T*tic term.
#pragma once
#include "d2ba_math.cuh"
#include "d2ba_types.cuh"

namespace d2ba {

// group-constant layout (doubles) inside the per-warp shared block
enum { GC_RA = 0, GC_RB = 9, GC_JW = 18, GC_JM = 27, GC_JC = 36, GC_RJI = 45, GC_TA = 54, GC_TB = 57, GC_TJI = 60, GC_WV = 63, GC_SIZE = 68 };

// Build the group constants.  R*/P* point at the rotation matrices / translations of the blocks the
// group touches (nullptr when absent).  Called by a full warp; `gc` is that warp's shared block.
D2BA_DEV void build_group_consts(int type, const double *Ri, const double *Pi, const double *Rj, const double *Pj,
                                 const double *Ra, const double *ta, const double *Rb, const double *tb, double *gc) {
  const int lane = threadIdx.x & 31;
  const bool one_frame = (type == P1F2C);
  const bool two_cam = (type == P2F2C || type == P1F2C);
  const double *Rb_ = two_cam ? Rb : Ra;
  const double *tb_ = two_cam ? tb : ta;
  if (lane < 9) { gc[GC_RA + lane] = Ra[lane]; gc[GC_RB + lane] = Rb_[lane]; }
  if (lane < 3) { gc[GC_TA + lane] = ta[lane]; gc[GC_TB + lane] = tb_[lane]; }
  // round 1: Jw = Rb^T Rj^T, Rji = Rj^T Ri, tji = Rj^T (Pi - Pj)
  if (lane < 9) {
    int i = lane / 3, j = lane % 3;
    double jw = 0, rji = (i == j) ? 1.0 : 0.0;
    if (!one_frame) {
      jw = Rb_[i] * Rj[j * 3] + Rb_[3 + i] * Rj[j * 3 + 1] + Rb_[6 + i] * Rj[j * 3 + 2];
      rji = Rj[i] * Ri[j] + Rj[3 + i] * Ri[3 + j] + Rj[6 + i] * Ri[6 + j];
    }
    gc[GC_JW + lane] = jw; gc[GC_RJI + lane] = rji;
  } else if (lane < 12) {
    int i = lane - 9;
    double t = 0;
    if (!one_frame) t = Rj[i] * (Pi[0] - Pj[0]) + Rj[3 + i] * (Pi[1] - Pj[1]) + Rj[6 + i] * (Pi[2] - Pj[2]);
    gc[GC_TJI + i] = t;
  }
  __syncwarp();
  // round 2: Jm = Jw Ri  (1F2C: Jm unused -> Rb^T so that Jc = Jm Ra = Rb^T Ra)
  if (lane < 9) {
    int i = lane / 3, j = lane % 3;
    double v;
    if (one_frame) v = Rb_[j * 3 + i];  // (Rb^T)[i][j]
    else v = gc[GC_JW + i * 3] * Ri[j] + gc[GC_JW + i * 3 + 1] * Ri[3 + j] + gc[GC_JW + i * 3 + 2] * Ri[6 + j];
    gc[GC_JM + lane] = v;
  }
  __syncwarp();
  // round 3: Jc = Jm Ra ; wv = Jw (Ri ta + Pi - Pj) - Rj^T ta   (reference: ...OneCamFactor.cpp:152-155)
  if (lane < 9) {
    int i = lane / 3, j = lane % 3;
    gc[GC_JC + lane] = gc[GC_JM + i * 3] * Ra[j] + gc[GC_JM + i * 3 + 1] * Ra[3 + j] + gc[GC_JM + i * 3 + 2] * Ra[6 + j];
  } else if (lane < 12) {
    int i = lane - 9;
    double v = 0;
    if (!one_frame) {
      double u[3];
#pragma unroll
      for (int k = 0; k < 3; k++) u[k] = Ri[k * 3] * ta[0] + Ri[k * 3 + 1] * ta[1] + Ri[k * 3 + 2] * ta[2] + Pi[k] - Pj[k];
      v = gc[GC_JW + i * 3] * u[0] + gc[GC_JW + i * 3 + 1] * u[1] + gc[GC_JW + i * 3 + 2] * u[2];
      v -= Rj[i] * ta[0] + Rj[3 + i] * ta[1] + Rj[6 + i] * ta[2];
    }
    gc[GC_WV + i] = v;
  }
  __syncwarp();
  if (one_frame && lane < 9) gc[GC_JM + lane] = 0.0;  // pose blocks do not exist for 1F2C
  __syncwarp();
}

template <int ROWS>
struct ProjOut {
  double r[ROWS];
  double J[4][ROWS][6];  // pose_i, pose_j, ext_a, ext_b
  double jl[ROWS], jt[ROWS];
  double cost;
};

// f: the kObsFields observation constants of this lane.  Output Jacobians are already
// loss-corrected (Huber, BaseParamResInfo.cpp:71-92: rho'' <= 0 so r, J scale by sqrt(rho')).
template <int ROWS, bool EXT, bool TD>
D2BA_DEV void proj_eval(int type, const double *gc, const double *f, double lam, double td, double s_px, double s_d,
                        double huber, ProjOut<ROWS> &o) {
  const double dti = td - f[12], dtj = td - f[13];
  double pi[3] = {f[0] - dti * f[6], f[1] - dti * f[7], f[2] - dti * f[8]};
  double pj[3] = {f[3] - dtj * f[9], f[4] - dtj * f[10], f[5] - dtj * f[11]};
  const double il = 1.0 / lam;
  double Pci[3] = {pi[0] * il, pi[1] * il, pi[2] * il};
  double Pmi[3], Pmj[3], Pcj[3], t[3];
  mv3(gc + GC_RA, Pci, Pmi);
  Pmi[0] += gc[GC_TA]; Pmi[1] += gc[GC_TA + 1]; Pmi[2] += gc[GC_TA + 2];
  mv3(gc + GC_RJI, Pmi, Pmj);
  Pmj[0] += gc[GC_TJI]; Pmj[1] += gc[GC_TJI + 1]; Pmj[2] += gc[GC_TJI + 2];
  t[0] = Pmj[0] - gc[GC_TB]; t[1] = Pmj[1] - gc[GC_TB + 1]; t[2] = Pmj[2] - gc[GC_TB + 2];
  mtv3(gc + GC_RB, t, Pcj);
  const double n2 = dot3(Pcj, Pcj), n = sqrt(n2), in = 1.0 / n;
  const double nj = sqrt(dot3(pj, pj)), inj = 1.0 / nj;
  double ph[3] = {Pcj[0] * in, Pcj[1] * in, Pcj[2] * in};
  double e[3] = {ph[0] - pj[0] * inj, ph[1] - pj[1] * inj, ph[2] - pj[2] * inj};
  const double *B = f + 14;  // tangent base rows b1, b2
  o.r[0] = s_px * dot3(B, e);
  o.r[1] = s_px * dot3(B + 3, e);
  // reduce = sqrt_info * tangent_base * (I/n - Pcj Pcj^T / n^3) = (s/n) (B - (B ph) ph^T)
  double red[ROWS][3];
  {
    double b0 = dot3(B, ph), b1 = dot3(B + 3, ph), sn = s_px * in;
#pragma unroll
    for (int k = 0; k < 3; k++) { red[0][k] = sn * (B[k] - b0 * ph[k]); red[1][k] = sn * (B[3 + k] - b1 * ph[k]); }
  }
  if (ROWS == 3) {
    o.r[2] = s_d * (in - f[20]);
    double c = -s_d * in * in;
#pragma unroll
    for (int k = 0; k < 3; k++) red[ROWS - 1][k] = c * ph[k];
  }
  // robust loss
  double s = 0;
#pragma unroll
  for (int q = 0; q < ROWS; q++) s += o.r[q] * o.r[q];
  double sc = 1.0;
  if (huber > 0 && s > huber * huber) {
    double rs = sqrt(s);
    o.cost = 0.5 * (2.0 * huber * rs - huber * huber);
    sc = sqrt(huber / rs);
  } else o.cost = 0.5 * s;
#pragma unroll
  for (int q = 0; q < ROWS; q++) {
    o.r[q] *= sc;
#pragma unroll
    for (int k = 0; k < 3; k++) red[q][k] *= sc;
  }
  double JcP[3];
  if (EXT) mv3(gc + GC_JC, Pci, JcP);
  double Ntv[3];
  if (TD) {
    // reduce_j_td * vel_j with the reference's mixed norms: I/|Pcj| - pj pj^T / |pj|^3
    double pv = dot3(pj, f + 9), i3 = inj * inj * inj;
#pragma unroll
    for (int k = 0; k < 3; k++) Ntv[k] = f[9 + k] * in - pj[k] * pv * i3;
  }
#pragma unroll
  for (int q = 0; q < ROWS; q++) {
    double A[3], Bm[3], Cr[3], Dc[3], c[3];
    rm3(red[q], gc + GC_JW, A);
    rm3(red[q], gc + GC_JM, Bm);
    rm3(red[q], gc + GC_JC, Dc);
    // Cr = red * Rb^T
    Cr[0] = red[q][0] * gc[GC_RB + 0] + red[q][1] * gc[GC_RB + 1] + red[q][2] * gc[GC_RB + 2];
    Cr[1] = red[q][0] * gc[GC_RB + 3] + red[q][1] * gc[GC_RB + 4] + red[q][2] * gc[GC_RB + 5];
    Cr[2] = red[q][0] * gc[GC_RB + 6] + red[q][1] * gc[GC_RB + 7] + red[q][2] * gc[GC_RB + 8];
    // pose_i : [ A , -(Bm x Pmi) ]
    cross3(Bm, Pmi, c);
    o.J[0][q][0] = A[0]; o.J[0][q][1] = A[1]; o.J[0][q][2] = A[2];
    o.J[0][q][3] = -c[0]; o.J[0][q][4] = -c[1]; o.J[0][q][5] = -c[2];
    // pose_j : [ -A , Cr x Pmj ]
    cross3(Cr, Pmj, c);
    o.J[1][q][0] = -A[0]; o.J[1][q][1] = -A[1]; o.J[1][q][2] = -A[2];
    o.J[1][q][3] = c[0]; o.J[1][q][4] = c[1]; o.J[1][q][5] = c[2];
    if (EXT) {
      double c2[3], c3[3];
      cross3(Dc, Pci, c);  // Dc x Pci
      if (type == P2F1C || type == P2F1CD) {
        cross3(red[q], JcP, c2);
        cross3(red[q], gc + GC_WV, c3);
#pragma unroll
        for (int k = 0; k < 3; k++) { o.J[2][q][k] = Bm[k] - Cr[k]; o.J[2][q][3 + k] = -c[k] + c2[k] + c3[k]; o.J[3][q][k] = 0; o.J[3][q][3 + k] = 0; }
      } else {
        cross3(red[q], Pcj, c2);
#pragma unroll
        for (int k = 0; k < 3; k++) {
          o.J[2][q][k] = (type == P1F2C) ? Cr[k] : Bm[k];
          o.J[2][q][3 + k] = -c[k];
          o.J[3][q][k] = -Cr[k];
          o.J[3][q][3 + k] = c2[k];
        }
      }
    }
    o.jl[q] = -il * dot3(Dc, Pci);
    if (TD) {
      double v = -il * dot3(Dc, f + 6);
      if (q < 2) v += sc * s_px * dot3(B + 3 * q, Ntv);
      o.jt[q] = v;
    }
  }
}

}  // namespace d2ba
