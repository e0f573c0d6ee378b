/*
 * orc_solver.c -- CPU ORACLE (test infrastructure): problem container, graph assembly,
 * Ceres-equivalent dense-Schur + traditional-dogleg trust-region solver, ADMM consensus loop.
 *
 * Follows (reference paths):
 *   graph assembly / factor-type dispatch  d2vins/src/estimator/d2estimator.cpp:758-877
 *   parameter-block lists per factor       d2vins/src/estimator/ParamResidualInfo.hpp:18-189
 *   solve driver                           d2common/src/solver/SolverWrapper.cpp:26-49
 *   ADMM outer loop                        d2common/src/solver/ConsensusSolver.cpp:39-235
 *   options                                d2vins/src/d2vins_params.cpp:140-160
 * The trust-region algorithm itself lives in un-vendored ceres-solver 2.1
 * (trust_region_minimizer.cc, dogleg_strategy.cc, schur_complement_solver.cc); it is restated
 * from its published algorithm (ASSUMED):
 *   - metric D = sqrt(diag(J^T J)) clamped to [1e-6, 1e32]; the minimizer's Jacobi scaling
 *     cancels analytically against the dogleg diagonal except inside that clamp, so the
 *     clamp is applied to the unscaled diagonal;
 *   - Gauss-Newton step from (J^T J + mu D^2) d = -g, mu from 1e-8, x10 on factorisation
 *     failure (max 1.0), relaxed to max(1e-8, mu/5) after success;
 *   - Cauchy step -alpha g~, alpha = |g~|^2/|J D^-1 g~|^2; traditional dogleg blend;
 *   - radius x0.5 when step quality < 0.25 or rejected, max(radius, 3|step|) when > 0.75;
 *   - step accepted iff cost decrease / model decrease > 1e-3; an attempt = one iteration.
 * The reduced-camera column order is: free POSE blocks (insertion order), free EXTRINSIC
 * blocks, TD if free [= landmark-coupled part, n_lc], then free SPEED_BIAS blocks [n_c].
 */
#define _GNU_SOURCE
#include "orc_oracle.h"
#include "orc_math.h"
#include <stdio.h>
#include <stdlib.h>
#include <time.h>
#include <pthread.h>

typedef struct { int type, pi, pj, ea, eb, lm, fa; orc_obs_const c; double depth; int64_t seq; } obs_t;   /* fa: anchor frame (relevance in marginalization) */
typedef struct { int pi, si, pj, sj; orc_imu_const c; } imu_t;

typedef struct { int64_t *keys; int *vals; int cap, n; } imap_t;
static void imap_init(imap_t *m) { m->cap = 0; m->n = 0; m->keys = NULL; m->vals = NULL; }
static void imap_free(imap_t *m) { free(m->keys); free(m->vals); imap_init(m); }
static uint64_t h64(uint64_t x) { x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33; return x; }
static int imap_get(const imap_t *m, int64_t k) {
  if (!m->cap) return -1;
  uint64_t i = h64((uint64_t)k) & (m->cap - 1);
  while (m->vals[i] >= 0) { if (m->keys[i] == k) return m->vals[i]; i = (i + 1) & (m->cap - 1); }
  return -1;
}
static void imap_put(imap_t *m, int64_t k, int v);
static void imap_grow(imap_t *m) {
  imap_t o = *m;
  m->cap = o.cap ? o.cap * 2 : 64; m->n = 0;
  m->keys = (int64_t *)malloc(sizeof(int64_t) * m->cap); m->vals = (int *)malloc(sizeof(int) * m->cap);
  for (int i = 0; i < m->cap; i++) m->vals[i] = -1;
  for (int i = 0; i < o.cap; i++) if (o.vals[i] >= 0) imap_put(m, o.keys[i], o.vals[i]);
  free(o.keys); free(o.vals);
}
static void imap_put(imap_t *m, int64_t k, int v) {
  if ((m->n + 1) * 2 > m->cap) imap_grow(m);
  uint64_t i = h64((uint64_t)k) & (m->cap - 1);
  while (m->vals[i] >= 0) { if (m->keys[i] == k) { m->vals[i] = v; return; } i = (i + 1) & (m->cap - 1); }
  m->keys[i] = k; m->vals[i] = v; m->n++;
}

typedef struct { double *pose, *ext, *sb, *lm; double td; } state_t;

typedef struct {
  double *Hcc, *gc, *hl, *gl, *W; /* W: nl x n_lc */
  double cost;
} lin_t;

struct orc_handle {
  d2ba_config cfg;
  int np, ne, nsb, nl, cap_p, cap_e, cap_s, cap_l;
  int64_t *pose_id, *ext_id, *sb_id, *lm_id;
  uint8_t *pose_c, *ext_c, *sb_c;
  uint8_t td_c; int has_td;
  imap_t pose_map, ext_map, sb_map, lm_map;
  state_t x, xc; /* current, candidate */
  obs_t *obs; int nobs, cap_obs;
  imu_t *imu; int nimu, cap_imu;
  /* prior */
  int pm, pnblk; double *pJ, *pe0, *px0; int *pkind, *pindex, *poff, *peff;
  /* consensus */
  int *pose_slot, *ext_slot; double *pose_z, *ext_z, *pose_tilde, *ext_tilde; int n_slots;
  int admm_on; double *lm_ref, *sb_ref; double td_ref;
  /* columns */
  int *pose_col, *ext_col, *sb_col; int td_col, n_lc, n_c, cols_valid;
  lin_t lin;
  double *S, *gred, *dc, *dl, *D2c, *D2l, *gn_c, *gn_l, *step_c, *step_l, *tmp_c, *tmp_l;
  double last_gn_valid;
  /* debug resjac */
  double *dbg_resjac;
};

static void *xrealloc(void *p, size_t n) { void *q = realloc(p, n ? n : 1); if (!q) abort(); return q; }
static double now_s(void) { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + 1e-9 * ts.tv_nsec; }

int orc_create(const d2ba_config *cfg, orc_handle **out) {
  orc_handle *o = (orc_handle *)calloc(1, sizeof(orc_handle));
  o->cfg = *cfg;
  if (o->cfg.initial_trust_region_radius <= 0) o->cfg.initial_trust_region_radius = 1e4;
  if (o->cfg.max_trust_region_radius <= 0) o->cfg.max_trust_region_radius = 1e16;
  if (o->cfg.min_relative_decrease <= 0) o->cfg.min_relative_decrease = 1e-3;
  if (o->cfg.function_tolerance <= 0) o->cfg.function_tolerance = 1e-6;
  if (o->cfg.gradient_tolerance <= 0) o->cfg.gradient_tolerance = 1e-10;
  if (o->cfg.parameter_tolerance <= 0) o->cfg.parameter_tolerance = 1e-8;
  imap_init(&o->pose_map); imap_init(&o->ext_map); imap_init(&o->sb_map); imap_init(&o->lm_map);
  o->td_c = 1; o->has_td = 0;
  *out = o;
  return 0;
}

static void free_lin(orc_handle *o) {
  free(o->lin.Hcc); free(o->lin.gc); free(o->lin.hl); free(o->lin.gl); free(o->lin.W);
  memset(&o->lin, 0, sizeof o->lin);
  free(o->S); free(o->gred); free(o->dc); free(o->dl); free(o->D2c); free(o->D2l); free(o->gn_c); free(o->gn_l);
  free(o->step_c); free(o->step_l); free(o->tmp_c); free(o->tmp_l);
  o->S = o->gred = o->dc = o->dl = o->D2c = o->D2l = o->gn_c = o->gn_l = o->step_c = o->step_l = o->tmp_c = o->tmp_l = NULL;
}

int orc_reset(orc_handle *o) {
  o->np = o->ne = o->nsb = o->nl = 0; o->nobs = o->nimu = 0; o->pm = o->pnblk = 0;
  imap_free(&o->pose_map); imap_free(&o->ext_map); imap_free(&o->sb_map); imap_free(&o->lm_map);
  o->has_td = 0; o->td_c = 1; o->admm_on = 0; o->n_slots = 0; o->cols_valid = 0;
  free_lin(o);
  return 0;
}

int orc_destroy(orc_handle *o) {
  if (!o) return 0;
  orc_reset(o);
  free(o->pose_id); free(o->ext_id); free(o->sb_id); free(o->lm_id);
  free(o->pose_c); free(o->ext_c); free(o->sb_c);
  free(o->x.pose); free(o->x.ext); free(o->x.sb); free(o->x.lm);
  free(o->xc.pose); free(o->xc.ext); free(o->xc.sb); free(o->xc.lm);
  free(o->obs); free(o->imu);
  free(o->pJ); free(o->pe0); free(o->px0); free(o->pkind); free(o->pindex); free(o->poff); free(o->peff);
  free(o->pose_slot); free(o->ext_slot); free(o->pose_z); free(o->ext_z); free(o->pose_tilde); free(o->ext_tilde);
  free(o->lm_ref); free(o->sb_ref);
  free(o->pose_col); free(o->ext_col); free(o->sb_col); free(o->dbg_resjac);
  free(o);
  return 0;
}

#define GROW(o, capf, n, ...)                                                    \
  if ((n) > (o)->capf) { int nc = (o)->capf ? (o)->capf * 2 : 64; while (nc < (n)) nc *= 2; (o)->capf = nc; __VA_ARGS__ }

int orc_set_blocks(orc_handle *o, int32_t kind, int32_t n, const int64_t *ids, const double *values,
                   const uint8_t *is_const) {
  o->cols_valid = 0;
  for (int i = 0; i < n; i++) {
    uint8_t c = is_const ? is_const[i] : 0;
    if (kind == D2BA_POSE) {
      int k = imap_get(&o->pose_map, ids[i]);
      if (k < 0) {
        k = o->np++;
        GROW(o, cap_p, o->np,
             o->pose_id = xrealloc(o->pose_id, sizeof(int64_t) * nc); o->pose_c = xrealloc(o->pose_c, nc);
             o->x.pose = xrealloc(o->x.pose, sizeof(double) * 7 * nc); o->xc.pose = xrealloc(o->xc.pose, sizeof(double) * 7 * nc);
             o->pose_slot = xrealloc(o->pose_slot, sizeof(int) * nc); o->pose_z = xrealloc(o->pose_z, sizeof(double) * 7 * nc);
             o->pose_tilde = xrealloc(o->pose_tilde, sizeof(double) * 6 * nc); o->pose_col = xrealloc(o->pose_col, sizeof(int) * nc);)
        o->pose_id[k] = ids[i]; o->pose_slot[k] = -1; imap_put(&o->pose_map, ids[i], k);
      }
      memcpy(o->x.pose + 7 * k, values + 7 * i, 7 * sizeof(double)); o->pose_c[k] = c;
    } else if (kind == D2BA_EXTRINSIC) {
      int k = imap_get(&o->ext_map, ids[i]);
      if (k < 0) {
        k = o->ne++;
        GROW(o, cap_e, o->ne,
             o->ext_id = xrealloc(o->ext_id, sizeof(int64_t) * nc); o->ext_c = xrealloc(o->ext_c, nc);
             o->x.ext = xrealloc(o->x.ext, sizeof(double) * 7 * nc); o->xc.ext = xrealloc(o->xc.ext, sizeof(double) * 7 * nc);
             o->ext_slot = xrealloc(o->ext_slot, sizeof(int) * nc); o->ext_z = xrealloc(o->ext_z, sizeof(double) * 7 * nc);
             o->ext_tilde = xrealloc(o->ext_tilde, sizeof(double) * 6 * nc); o->ext_col = xrealloc(o->ext_col, sizeof(int) * nc);)
        o->ext_id[k] = ids[i]; o->ext_slot[k] = -1; imap_put(&o->ext_map, ids[i], k);
      }
      memcpy(o->x.ext + 7 * k, values + 7 * i, 7 * sizeof(double)); o->ext_c[k] = c;
    } else if (kind == D2BA_SPEED_BIAS) {
      int k = imap_get(&o->sb_map, ids[i]);
      if (k < 0) {
        k = o->nsb++;
        GROW(o, cap_s, o->nsb,
             o->sb_id = xrealloc(o->sb_id, sizeof(int64_t) * nc); o->sb_c = xrealloc(o->sb_c, nc);
             o->x.sb = xrealloc(o->x.sb, sizeof(double) * 9 * nc); o->xc.sb = xrealloc(o->xc.sb, sizeof(double) * 9 * nc);
             o->sb_ref = xrealloc(o->sb_ref, sizeof(double) * 9 * nc); o->sb_col = xrealloc(o->sb_col, sizeof(int) * nc);)
        o->sb_id[k] = ids[i]; imap_put(&o->sb_map, ids[i], k);
      }
      memcpy(o->x.sb + 9 * k, values + 9 * i, 9 * sizeof(double)); o->sb_c[k] = c;
    } else if (kind == D2BA_TD) {
      o->x.td = values[i]; o->td_c = c; o->has_td = 1;
    } else if (kind == D2BA_LANDMARK) {
      int k = imap_get(&o->lm_map, ids[i]);
      if (k < 0) {
        k = o->nl++;
        GROW(o, cap_l, o->nl,
             o->lm_id = xrealloc(o->lm_id, sizeof(int64_t) * nc);
             o->x.lm = xrealloc(o->x.lm, sizeof(double) * nc); o->xc.lm = xrealloc(o->xc.lm, sizeof(double) * nc);
             o->lm_ref = xrealloc(o->lm_ref, sizeof(double) * nc);)
        o->lm_id[k] = ids[i]; imap_put(&o->lm_map, ids[i], k);
      }
      o->x.lm[k] = values[i];
    } else return 1;
  }
  return 0;
}

int orc_add_proj(orc_handle *o, int32_t n, const d2ba_proj_obs *in) {
  for (int i = 0; i < n; i++) {
    const d2ba_proj_obs *p = in + i;
    obs_t t;
    memset(&t, 0, sizeof t);
    t.type = p->type; t.pi = t.pj = t.ea = t.eb = -1;
    t.lm = imap_get(&o->lm_map, p->landmark_id);
    if (t.lm < 0) return 2;
    t.fa = imap_get(&o->pose_map, p->frame_a);
    if (p->type != D2BA_PROJ_DEPTH_PRIOR) {
      /* parameter lists: ParamResidualInfo.hpp:34-43 (2F1C), :72-82 (2F2C), :107-115 (1F2C) */
      t.ea = imap_get(&o->ext_map, p->cam_a);
      if (t.ea < 0) return 3;
      if (p->type == D2BA_PROJ_2F2C || p->type == D2BA_PROJ_1F2C) { t.eb = imap_get(&o->ext_map, p->cam_b); if (t.eb < 0) return 3; }
      if (p->type != D2BA_PROJ_1F2C) {
        t.pi = imap_get(&o->pose_map, p->frame_a); t.pj = imap_get(&o->pose_map, p->frame_b);
        if (t.pi < 0 || t.pj < 0) return 4;
      }
      memcpy(t.c.pts_i, p->pts_i, 24); memcpy(t.c.pts_j, p->pts_j, 24);
      memcpy(t.c.vel_i, p->vel_i, 24); memcpy(t.c.vel_j, p->vel_j, 24);
      t.c.td_i = p->td_i; t.c.td_j = p->td_j;
      orc_tangent_base(p->pts_j, t.c.tangent_base);
      if (p->type == D2BA_PROJ_2F1C_DEPTH) t.c.inv_depth_j = 1.0 / p->depth;
    }
    t.depth = p->depth;
    t.seq = o->nobs;
    o->nobs++;
    GROW(o, cap_obs, o->nobs, o->obs = xrealloc(o->obs, sizeof(obs_t) * nc);)
    o->obs[o->nobs - 1] = t;
  }
  return 0;
}

/* D2Estimator::setupLandmarkFactors dispatch, d2vins/src/estimator/d2estimator.cpp:796-874:
 * anchor = track[0]; optional OneFrameDepth on the anchor (:806-815); for track[k>=1]:
 * same camera -> 2F1C (Depth variant when valid & fuse_dep), skipped if same frame (:839-846);
 * other camera & same frame -> 1F2C; other camera & other frame -> 2F2C. */
int orc_add_landmark_tracks(orc_handle *o, int32_t n_landmarks, const int64_t *landmark_ids,
                            const int32_t *track_ptr, const d2ba_track_obs *tobs, int32_t fuse_dep,
                            double min_d, double max_d, int32_t n_ignore, const int64_t *ignore) {
  for (int l = 0; l < n_landmarks; l++) {
    int b = track_ptr[l], e = track_ptr[l + 1];
    if (e - b < 1) continue;
    const d2ba_track_obs *first = tobs + b;
    int ign = 0;
    for (int k = 0; k < n_ignore; k++) if (ignore[k] == first->frame_id) ign = 1;
    if (ign) continue;
    d2ba_proj_obs p;
    if (first->depth_mea && fuse_dep && first->depth < max_d && first->depth > min_d) {
      memset(&p, 0, sizeof p);
      p.type = D2BA_PROJ_DEPTH_PRIOR; p.frame_a = first->frame_id; p.landmark_id = landmark_ids[l]; p.depth = first->depth;
      p.cam_a = first->camera_id;
      int rc = orc_add_proj(o, 1, &p); if (rc) return rc;
    }
    for (int k = b + 1; k < e; k++) {
      const d2ba_track_obs *t = tobs + k;
      ign = 0;
      for (int q = 0; q < n_ignore; q++) if (ignore[q] == t->frame_id) ign = 1;
      if (ign) continue;
      memset(&p, 0, sizeof p);
      p.frame_a = first->frame_id; p.frame_b = t->frame_id; p.landmark_id = landmark_ids[l];
      p.cam_a = first->camera_id; p.cam_b = t->camera_id;
      memcpy(p.pts_i, first->pt3d_norm, 24); memcpy(p.pts_j, t->pt3d_norm, 24);
      memcpy(p.vel_i, first->velocity, 24); memcpy(p.vel_j, t->velocity, 24);
      p.td_i = first->cur_td; p.td_j = t->cur_td;
      if (t->camera_id == first->camera_id) {
        if (t->depth_mea && fuse_dep && t->depth < max_d && t->depth > min_d) { p.type = D2BA_PROJ_2F1C_DEPTH; p.depth = t->depth; }
        else p.type = D2BA_PROJ_2F1C;
        if (first->frame_id == t->frame_id) continue;
      } else if (t->frame_id == first->frame_id) p.type = D2BA_PROJ_1F2C;
      else p.type = D2BA_PROJ_2F2C;
      int rc = orc_add_proj(o, 1, &p); if (rc) return rc;
    }
  }
  return 0;
}

int orc_add_imu(orc_handle *o, int32_t n, const d2ba_imu *in) {
  for (int i = 0; i < n; i++) {
    imu_t t;
    /* ImuResInfo::paramsList, ParamResidualInfo.hpp:134-142 */
    t.pi = imap_get(&o->pose_map, in[i].frame_a); t.pj = imap_get(&o->pose_map, in[i].frame_b);
    t.si = imap_get(&o->sb_map, in[i].frame_a); t.sj = imap_get(&o->sb_map, in[i].frame_b);
    if (t.pi < 0 || t.pj < 0 || t.si < 0 || t.sj < 0) return 5;
    t.c.sum_dt = in[i].sum_dt;
    memcpy(t.c.delta_p, in[i].delta_p, 24); memcpy(t.c.delta_q, in[i].delta_q, 32); memcpy(t.c.delta_v, in[i].delta_v, 24);
    memcpy(t.c.linearized_ba, in[i].linearized_ba, 24); memcpy(t.c.linearized_bg, in[i].linearized_bg, 24);
    memcpy(t.c.jacobian, in[i].jacobian, sizeof t.c.jacobian); memcpy(t.c.covariance, in[i].covariance, sizeof t.c.covariance);
    if (orc_imu_sqrt_info(t.c.covariance, t.c.sqrt_info)) return 6;
    o->nimu++;
    GROW(o, cap_imu, o->nimu, o->imu = xrealloc(o->imu, sizeof(imu_t) * nc);)
    o->imu[o->nimu - 1] = t;
  }
  return 0;
}

static int kind_size(int kind) { return kind == D2BA_POSE || kind == D2BA_EXTRINSIC ? 7 : kind == D2BA_SPEED_BIAS ? 9 : 1; }
static int kind_eff(int kind) { return kind == D2BA_POSE || kind == D2BA_EXTRINSIC ? 6 : kind == D2BA_SPEED_BIAS ? 9 : 1; }

static int find_block(orc_handle *o, int kind, int64_t id) {
  switch (kind) {
    case D2BA_POSE: return imap_get(&o->pose_map, id);
    case D2BA_EXTRINSIC: return imap_get(&o->ext_map, id);
    case D2BA_SPEED_BIAS: return imap_get(&o->sb_map, id);
    case D2BA_TD: return o->has_td ? 0 : -1;
    case D2BA_LANDMARK: return imap_get(&o->lm_map, id);
  }
  return -1;
}

int orc_set_prior(orc_handle *o, int32_t m, const double *J, const double *e0, int32_t nblk,
                  const d2ba_blockref *refs, const double *x0) {
  o->pm = m; o->pnblk = nblk;
  o->pJ = xrealloc(o->pJ, sizeof(double) * m * m); memcpy(o->pJ, J, sizeof(double) * m * m);
  o->pe0 = xrealloc(o->pe0, sizeof(double) * m); memcpy(o->pe0, e0, sizeof(double) * m);
  o->pkind = xrealloc(o->pkind, sizeof(int) * nblk); o->pindex = xrealloc(o->pindex, sizeof(int) * nblk);
  o->poff = xrealloc(o->poff, sizeof(int) * nblk); o->peff = xrealloc(o->peff, sizeof(int) * nblk);
  o->px0 = xrealloc(o->px0, sizeof(double) * 9 * nblk);
  int off = 0, xo = 0;
  for (int i = 0; i < nblk; i++) {
    o->pkind[i] = refs[i].kind; o->pindex[i] = find_block(o, refs[i].kind, refs[i].id);
    if (o->pindex[i] < 0) return 7;
    o->poff[i] = off; o->peff[i] = kind_eff(refs[i].kind); off += o->peff[i];
    memcpy(o->px0 + 9 * i, x0 + xo, sizeof(double) * kind_size(refs[i].kind)); xo += kind_size(refs[i].kind);
  }
  if (off != m) return 8;
  return 0;
}

int orc_set_prior_info(orc_handle *o, int32_t m, const double *A, const double *b, int32_t nblk,
                       const d2ba_blockref *refs, const double *x0) {
  double *J = (double *)malloc(sizeof(double) * m * m), *e0 = (double *)malloc(sizeof(double) * m);
  orc_to_jac_res(m, A, b, J, e0);
  int rc = orc_set_prior(o, m, J, e0, nblk, refs, x0);
  free(J); free(e0);
  return rc;
}

int orc_set_consensus(orc_handle *o, int32_t n, const d2ba_blockref *refs, const int32_t *slot,
                      int32_t n_slots_global) {
  o->admm_on = 1; o->n_slots = n_slots_global;
  for (int i = 0; i < n; i++) {
    int k = find_block(o, refs[i].kind, refs[i].id);
    if (k < 0) return 9;
    if (refs[i].kind == D2BA_POSE) o->pose_slot[k] = slot[i];
    else if (refs[i].kind == D2BA_EXTRINSIC) o->ext_slot[k] = slot[i];
    else return 10;
  }
  return 0;
}

/* ------------------------------------------------------------------ columns */
static void assign_cols(orc_handle *o) {
  int c = 0;
  for (int i = 0; i < o->np; i++) { o->pose_col[i] = o->pose_c[i] ? -1 : c; if (!o->pose_c[i]) c += 6; }
  for (int i = 0; i < o->ne; i++) { o->ext_col[i] = o->ext_c[i] ? -1 : c; if (!o->ext_c[i]) c += 6; }
  o->td_col = (o->has_td && !o->td_c) ? c : -1;
  if (o->td_col >= 0) c += 1;
  o->n_lc = c;
  for (int i = 0; i < o->nsb; i++) { o->sb_col[i] = o->sb_c[i] ? -1 : c; if (!o->sb_c[i]) c += 9; }
  o->n_c = c;
  free_lin(o);
  size_t nc = o->n_c ? o->n_c : 1, nl = o->nl ? o->nl : 1, nlc = o->n_lc ? o->n_lc : 1;
  o->lin.Hcc = calloc(nc * nc, 8); o->lin.gc = calloc(nc, 8); o->lin.hl = calloc(nl, 8); o->lin.gl = calloc(nl, 8);
  o->lin.W = calloc(nl * nlc, 8);
  o->S = calloc(nc * nc, 8); o->gred = calloc(nc, 8); o->dc = calloc(nc, 8); o->dl = calloc(nl, 8);
  o->D2c = calloc(nc, 8); o->D2l = calloc(nl, 8); o->gn_c = calloc(nc, 8); o->gn_l = calloc(nl, 8);
  o->step_c = calloc(nc, 8); o->step_l = calloc(nl, 8); o->tmp_c = calloc(nc, 8); o->tmp_l = calloc(nl, 8);
  o->cols_valid = 1;
}

static void copy_state(orc_handle *o, state_t *d, const state_t *s) {
  memcpy(d->pose, s->pose, sizeof(double) * 7 * o->np); memcpy(d->ext, s->ext, sizeof(double) * 7 * o->ne);
  memcpy(d->sb, s->sb, sizeof(double) * 9 * o->nsb); memcpy(d->lm, s->lm, sizeof(double) * o->nl);
  d->td = s->td;
}

/* ------------------------------------------------------------------ accumulation */
typedef struct { int col, width, ld; const double *J; } jblk_t;

static void accumulate(orc_handle *o, lin_t *L, int rows, int nb, const jblk_t *b, const double *r, int lm,
                       const double *Jl) {
  int n = o->n_c;
  for (int a = 0; a < nb; a++) {
    if (b[a].col < 0) continue;
    for (int i = 0; i < b[a].width; i++) {
      double g = 0;
      for (int q = 0; q < rows; q++) g += b[a].J[q * b[a].ld + i] * r[q];
      L->gc[b[a].col + i] += g;
    }
    for (int c = 0; c < nb; c++) {
      if (b[c].col < 0) continue;
      for (int i = 0; i < b[a].width; i++)
        for (int j = 0; j < b[c].width; j++) {
          double s = 0;
          for (int q = 0; q < rows; q++) s += b[a].J[q * b[a].ld + i] * b[c].J[q * b[c].ld + j];
          L->Hcc[(size_t)(b[a].col + i) * n + b[c].col + j] += s;
        }
    }
    if (lm >= 0)
      for (int i = 0; i < b[a].width; i++) {
        double s = 0;
        for (int q = 0; q < rows; q++) s += b[a].J[q * b[a].ld + i] * Jl[q];
        L->W[(size_t)lm * o->n_lc + b[a].col + i] += s;
      }
  }
  if (lm >= 0) {
    double h = 0, g = 0;
    for (int q = 0; q < rows; q++) { h += Jl[q] * Jl[q]; g += Jl[q] * r[q]; }
    L->hl[lm] += h; L->gl[lm] += g;
  }
}

/* evaluate one projection residual with loss correction; returns cost contribution */
static double eval_proj(orc_handle *o, const state_t *x, const obs_t *t, int want_jac, double *r,
                        double *Ji, double *Jj, double *Ja, double *Jb, double *Jl, double *Jtd, int *rows_out) {
  double sq = o->cfg.focal_length / 1.5;
  int rows;
  if (t->type == ORC_PROJ_DEPTH_PRIOR) {
    rows = 1;
    orc_depth_prior_eval(x->lm[t->lm], t->depth, o->cfg.depth_sqrt_inf, r, want_jac ? Jl : NULL);
  } else {
    rows = t->type == ORC_PROJ_2F1C_DEPTH ? 3 : 2;
    orc_proj_eval(t->type, &t->c, sq, o->cfg.depth_sqrt_inf, t->pi >= 0 ? x->pose + 7 * t->pi : NULL,
                  t->pj >= 0 ? x->pose + 7 * t->pj : NULL, x->ext + 7 * t->ea, t->eb >= 0 ? x->ext + 7 * t->eb : NULL,
                  x->lm[t->lm], x->td, r, want_jac ? Ji : NULL, want_jac ? Jj : NULL, want_jac ? Ja : NULL,
                  want_jac ? Jb : NULL, want_jac ? Jl : NULL, want_jac ? Jtd : NULL);
  }
  *rows_out = rows;
  double s = 0;
  for (int q = 0; q < rows; q++) s += r[q] * r[q];
  if (o->cfg.huber_delta > 0) {
    /* loss applied to every landmark residual incl. depth (d2estimator.cpp:764) */
    double rho[3], rs, sr1, asn;
    orc_huber(o->cfg.huber_delta, s, rho);
    orc_corrector(rho, s, &rs, &sr1, &asn);
    if (want_jac) {
      /* J <- sqrt(rho1) (J - alpha_sq_norm r (r^T J)); alpha_sq_norm == 0 for Huber */
      double *Js[6] = {Ji, Jj, Ja, Jb, Jl, Jtd};
      int w[6] = {7, 7, 7, 7, 1, 1};
      for (int k = 0; k < 6; k++) {
        if (t->type == ORC_PROJ_DEPTH_PRIOR && k != 4) continue;
        for (int c = 0; c < w[k]; c++) {
          double rtj = 0;
          for (int q = 0; q < rows; q++) rtj += r[q] * Js[k][q * w[k] + c];
          for (int q = 0; q < rows; q++) Js[k][q * w[k] + c] = sr1 * (Js[k][q * w[k] + c] - asn * r[q] * rtj);
        }
      }
    }
    for (int q = 0; q < rows; q++) r[q] *= rs;
    return 0.5 * rho[0];
  }
  return 0.5 * s;
}

static double linearize(orc_handle *o, const state_t *x, lin_t *L, int want_jac) {
  int n = o->n_c;
  double cost = 0;
  if (want_jac) {
    memset(L->Hcc, 0, sizeof(double) * (size_t)n * n); memset(L->gc, 0, sizeof(double) * n);
    memset(L->hl, 0, sizeof(double) * o->nl); memset(L->gl, 0, sizeof(double) * o->nl);
    memset(L->W, 0, sizeof(double) * (size_t)o->nl * o->n_lc);
  }
  /* reprojection */
  for (int k = 0; k < o->nobs; k++) {
    const obs_t *t = o->obs + k;
    double r[3], Ji[21] = {0}, Jj[21] = {0}, Ja[21] = {0}, Jb[21] = {0}, Jl[3] = {0}, Jtd[3] = {0};
    int rows;
    cost += eval_proj(o, x, t, want_jac, r, Ji, Jj, Ja, Jb, Jl, Jtd, &rows);
    if (!want_jac) continue;
    jblk_t b[5]; int nb = 0;
    if (t->type != ORC_PROJ_DEPTH_PRIOR) {
      if (t->pi >= 0) { b[nb++] = (jblk_t){o->pose_col[t->pi], 6, 7, Ji}; b[nb++] = (jblk_t){o->pose_col[t->pj], 6, 7, Jj}; }
      b[nb++] = (jblk_t){o->ext_col[t->ea], 6, 7, Ja};
      if (t->eb >= 0) b[nb++] = (jblk_t){o->ext_col[t->eb], 6, 7, Jb};
      b[nb++] = (jblk_t){o->td_col, 1, 1, Jtd};
    }
    accumulate(o, L, rows, nb, b, r, t->lm, Jl);
  }
  /* IMU */
  for (int k = 0; k < o->nimu; k++) {
    const imu_t *t = o->imu + k;
    double r[15], Jpi[105], Jsi[135], Jpj[105], Jsj[135];
    orc_imu_eval(&t->c, o->cfg.gravity_norm, x->pose + 7 * t->pi, x->sb + 9 * t->si, x->pose + 7 * t->pj,
                 x->sb + 9 * t->sj, r, want_jac ? Jpi : NULL, want_jac ? Jsi : NULL, want_jac ? Jpj : NULL, want_jac ? Jsj : NULL);
    double s = 0; for (int q = 0; q < 15; q++) s += r[q] * r[q];
    cost += 0.5 * s;
    if (!want_jac) continue;
    jblk_t b[4] = {{o->pose_col[t->pi], 6, 7, Jpi}, {o->sb_col[t->si], 9, 9, Jsi}, {o->pose_col[t->pj], 6, 7, Jpj}, {o->sb_col[t->sj], 9, 9, Jsj}};
    accumulate(o, L, 15, 4, b, r, -1, NULL);
  }
  /* prior: r = e0 + J dx  (prior_factor.cpp:45-90) */
  if (o->pm > 0) {
    int m = o->pm;
    double *dx = (double *)malloc(sizeof(double) * m), *r = (double *)malloc(sizeof(double) * m);
    for (int i = 0; i < o->pnblk; i++) {
      int kind = o->pkind[i], idx = o->pindex[i], off = o->poff[i];
      const double *x0 = o->px0 + 9 * i;
      if (kind == D2BA_POSE) orc_prior_dx_pose(x->pose + 7 * idx, x0, dx + off);
      else if (kind == D2BA_EXTRINSIC) orc_prior_dx_pose(x->ext + 7 * idx, x0, dx + off);
      else if (kind == D2BA_SPEED_BIAS) for (int q = 0; q < 9; q++) dx[off + q] = x->sb[9 * idx + q] - x0[q];
      else if (kind == D2BA_TD) dx[off] = x->td - x0[0];
      else dx[off] = x->lm[idx] - x0[0];
    }
    double s = 0;
    for (int i = 0; i < m; i++) {
      double a = o->pe0[i];
      for (int j = 0; j < m; j++) a += o->pJ[(size_t)i * m + j] * dx[j];
      r[i] = a; s += a * a;
    }
    cost += 0.5 * s;
    if (want_jac) {
      jblk_t *b = (jblk_t *)malloc(sizeof(jblk_t) * o->pnblk);
      int lm = -1; const double *Jl = NULL; double *Jlbuf = NULL;
      int nb = 0;
      for (int i = 0; i < o->pnblk; i++) {
        int kind = o->pkind[i], idx = o->pindex[i];
        int col = kind == D2BA_POSE ? o->pose_col[idx] : kind == D2BA_EXTRINSIC ? o->ext_col[idx] : kind == D2BA_SPEED_BIAS ? o->sb_col[idx] : kind == D2BA_TD ? o->td_col : -2;
        if (col == -2) { /* landmark inside a prior: not produced by the reference's marginalizer (remove_base_when_margin_remote=2) */
          lm = idx; Jlbuf = (double *)malloc(sizeof(double) * m);
          for (int q = 0; q < m; q++) Jlbuf[q] = o->pJ[(size_t)q * m + o->poff[i]];
          Jl = Jlbuf; continue;
        }
        b[nb++] = (jblk_t){col, o->peff[i], m, o->pJ + o->poff[i]};
      }
      accumulate(o, L, m, nb, b, r, lm, Jl);
      free(b); free(Jlbuf);
    }
    free(dx); free(r);
  }
  /* ADMM terms (ConsensusSolver::updateTilde, ConsensusSolver.cpp:108-164) */
  if (o->admm_on) {
    for (int pass = 0; pass < 2; pass++) {
      int nblk = pass == 0 ? o->np : o->ne;
      for (int i = 0; i < nblk; i++) {
        int slot = pass == 0 ? o->pose_slot[i] : o->ext_slot[i];
        if (slot < 0) continue;
        const double *z = (pass == 0 ? o->pose_z : o->ext_z) + 7 * i, *tl = (pass == 0 ? o->pose_tilde : o->ext_tilde) + 6 * i;
        const double *xp = (pass == 0 ? x->pose : x->ext) + 7 * i;
        double r[6], J[42];
        orc_consensus_eval(z, z + 3, tl, tl + 3, o->cfg.rho_frame_T, o->cfg.rho_frame_theta, xp, r, want_jac ? J : NULL);
        double s = 0; for (int q = 0; q < 6; q++) s += r[q] * r[q];
        cost += 0.5 * s;
        if (want_jac) { jblk_t b = {pass == 0 ? o->pose_col[i] : o->ext_col[i], 6, 7, J}; accumulate(o, L, 6, 1, &b, r, -1, NULL); }
      }
    }
    /* ceres::NormalPrior(A, x_ref) on local-only params: A = rho_landmark*I for LANDMARK, I otherwise (:113-125) */
    for (int l = 0; l < o->nl; l++) {
      double A = o->cfg.rho_landmark, r = A * (x->lm[l] - o->lm_ref[l]);
      cost += 0.5 * r * r;
      if (want_jac) { L->hl[l] += A * A; L->gl[l] += A * r; }
    }
    for (int i = 0; i < o->nsb; i++) {
      double r[9], J[81] = {0};
      for (int q = 0; q < 9; q++) { r[q] = x->sb[9 * i + q] - o->sb_ref[9 * i + q]; J[q * 9 + q] = 1.0; cost += 0.5 * r[q] * r[q]; }
      if (want_jac) { jblk_t b = {o->sb_col[i], 9, 9, J}; accumulate(o, L, 9, 1, &b, r, -1, NULL); }
    }
    if (o->has_td) {
      double r = x->td - o->td_ref, J = 1.0;
      cost += 0.5 * r * r;
      if (want_jac) { jblk_t b = {o->td_col, 1, 1, &J}; accumulate(o, L, 1, 1, &b, &r, -1, NULL); }
    }
  }
  if (want_jac) L->cost = cost;
  return cost;
}

/* ------------------------------------------------------------------ linear algebra */
static int chol_inplace(double *A, int n) { /* lower, row-major; returns 0 ok */
  for (int i = 0; i < n; i++) {
    double *Ai = A + (size_t)i * n;
    for (int j = 0; j <= i; j++) {
      const double *Aj = A + (size_t)j * n;
      double s = Ai[j];
      for (int k = 0; k < j; k++) s -= Ai[k] * Aj[k];
      if (i == j) { if (!(s > 0) || !isfinite(s)) return 1; Ai[j] = sqrt(s); }
      else Ai[j] = s / Aj[j];
    }
  }
  return 0;
}
static void chol_solve(const double *Lm, int n, double *b) {
  for (int i = 0; i < n; i++) { double s = b[i]; for (int k = 0; k < i; k++) s -= Lm[(size_t)i * n + k] * b[k]; b[i] = s / Lm[(size_t)i * n + i]; }
  for (int i = n - 1; i >= 0; i--) { double s = b[i]; for (int k = i + 1; k < n; k++) s -= Lm[(size_t)k * n + i] * b[k]; b[i] = s / Lm[(size_t)i * n + i]; }
}

/* Gauss-Newton step of (H + mu D^2) d = -g by landmark Schur elimination + dense Cholesky
 * (the arithmetic ceres DENSE_SCHUR performs). returns 0 ok */
static int solve_gn(orc_handle *o, const lin_t *L, double mu, double *dc, double *dl, int keep_S) {
  int n = o->n_c, nlc = o->n_lc;
  double *S = o->S, *gr = o->gred;
  memcpy(S, L->Hcc, sizeof(double) * (size_t)n * n); memcpy(gr, L->gc, sizeof(double) * n);
  for (int i = 0; i < n; i++) S[(size_t)i * n + i] += mu * o->D2c[i];
  int *nz = (int *)malloc(sizeof(int) * (nlc + 1));
  for (int l = 0; l < o->nl; l++) {
    double h = L->hl[l] + mu * o->D2l[l];
    if (!(h > 0)) { free(nz); return 1; }
    const double *w = L->W + (size_t)l * nlc;
    int cnt = 0;
    for (int i = 0; i < nlc; i++) if (w[i] != 0.0) nz[cnt++] = i;
    double ih = 1.0 / h, gl = L->gl[l];
    for (int a = 0; a < cnt; a++) {
      int i = nz[a]; double wi = w[i] * ih;
      gr[i] -= wi * gl;
      double *Si = S + (size_t)i * n;
      for (int b = 0; b <= a; b++) Si[nz[b]] -= wi * w[nz[b]];
    }
  }
  free(nz);
  /* mirror lower -> upper for debug */
  if (keep_S) for (int i = 0; i < n; i++) for (int j = 0; j < i; j++) S[(size_t)j * n + i] = S[(size_t)i * n + j];
  if (keep_S == 2) return 0;
  if (n > 0 && chol_inplace(S, n)) return 2;
  for (int i = 0; i < n; i++) dc[i] = -gr[i];
  if (n > 0) chol_solve(S, n, dc);
  for (int l = 0; l < o->nl; l++) {
    double h = L->hl[l] + mu * o->D2l[l];
    const double *w = L->W + (size_t)l * nlc;
    double s = L->gl[l];
    for (int i = 0; i < nlc; i++) s += w[i] * dc[i];
    dl[l] = -s / h;
  }
  for (int i = 0; i < n; i++) if (!isfinite(dc[i])) return 3;
  return 0;
}

/* y = H x with H in Schur-block form */
static void H_mul(orc_handle *o, const lin_t *L, const double *xc, const double *xl, double *yc, double *yl) {
  int n = o->n_c, nlc = o->n_lc;
  for (int i = 0; i < n; i++) { double s = 0; const double *Hi = L->Hcc + (size_t)i * n; for (int j = 0; j < n; j++) s += Hi[j] * xc[j]; yc[i] = s; }
  for (int l = 0; l < o->nl; l++) {
    const double *w = L->W + (size_t)l * nlc;
    double s = 0;
    for (int i = 0; i < nlc; i++) { s += w[i] * xc[i]; yc[i] += w[i] * xl[l]; }
    yl[l] = s + L->hl[l] * xl[l];
  }
}

static void apply_step(orc_handle *o, const state_t *x, const double *sc, const double *sl, state_t *out) {
  copy_state(o, out, x);
  for (int i = 0; i < o->np; i++) if (o->pose_col[i] >= 0) orc_pose_plus(x->pose + 7 * i, sc + o->pose_col[i], out->pose + 7 * i);
  for (int i = 0; i < o->ne; i++) if (o->ext_col[i] >= 0) orc_pose_plus(x->ext + 7 * i, sc + o->ext_col[i], out->ext + 7 * i);
  if (o->td_col >= 0) out->td = x->td + sc[o->td_col];
  for (int i = 0; i < o->nsb; i++) if (o->sb_col[i] >= 0) for (int q = 0; q < 9; q++) out->sb[9 * i + q] = x->sb[9 * i + q] + sc[o->sb_col[i] + q];
  for (int l = 0; l < o->nl; l++) out->lm[l] = x->lm[l] + sl[l];
}

static void norms(orc_handle *o, const state_t *x, const state_t *c, double *x_norm, double *step_norm) {
  double xn = 0, sn = 0;
#define ACC(ptr, cptr, k) { double d = (ptr)[k] - (cptr)[k]; xn += (ptr)[k] * (ptr)[k]; sn += d * d; }
  for (int i = 0; i < o->np; i++) if (o->pose_col[i] >= 0) for (int q = 0; q < 7; q++) ACC(x->pose + 7 * i, c->pose + 7 * i, q)
  for (int i = 0; i < o->ne; i++) if (o->ext_col[i] >= 0) for (int q = 0; q < 7; q++) ACC(x->ext + 7 * i, c->ext + 7 * i, q)
  if (o->td_col >= 0) { double d = x->td - c->td; xn += x->td * x->td; sn += d * d; }
  for (int i = 0; i < o->nsb; i++) if (o->sb_col[i] >= 0) for (int q = 0; q < 9; q++) ACC(x->sb + 9 * i, c->sb + 9 * i, q)
  for (int l = 0; l < o->nl; l++) ACC(x->lm, c->lm, l)
#undef ACC
  *x_norm = sqrt(xn); *step_norm = sqrt(sn);
}

static double grad_max(orc_handle *o, const lin_t *L) {
  double m = 0;
  for (int i = 0; i < o->n_c; i++) if (fabs(L->gc[i]) > m) m = fabs(L->gc[i]);
  for (int l = 0; l < o->nl; l++) if (fabs(L->gl[l]) > m) m = fabs(L->gl[l]);
  return m;
}

/* one ceres::Solve: trust-region loop on the current problem. fixed>0: run exactly `max_iter`
 * attempts with convergence exits disabled. */
static void tr_solve(orc_handle *o, int max_iter, int fixed, d2ba_report *rep) {
  const d2ba_config *cf = &o->cfg;
  double t0 = now_s();
  if (!o->cols_valid) assign_cols(o);
  lin_t *L = &o->lin;
  int n = o->n_c, nl = o->nl;
  double cost = linearize(o, &o->x, L, 1);
  rep->initial_cost = cost; rep->total_iterations = 0; rep->successful_steps = 0; rep->succ = 1;
  rep->termination = D2BA_TERM_NO_CONVERGENCE;
  double radius = cf->initial_trust_region_radius, mu = 1e-8;
  const double min_mu = 1e-8, max_mu = 1.0, mu_inc = 10.0;
  int reuse = 0, invalid_run = 0;
  double gg = 0, nn = 0, gdn = 0, alpha = 0;
  double gmax = grad_max(o, L);
  if (!fixed && gmax <= cf->gradient_tolerance) { rep->termination = D2BA_TERM_GRADIENT_TOL; goto done; }
  for (int it = 0; it < max_iter; it++) {
    if (!reuse) {
      /* trust-region metric */
      for (int i = 0; i < n; i++) { double d = sqrt(L->Hcc[(size_t)i * n + i]); d = d < 1e-6 ? 1e-6 : (d > 1e32 ? 1e32 : d); o->D2c[i] = d * d; }
      for (int l = 0; l < nl; l++) { double d = sqrt(L->hl[l]); d = d < 1e-6 ? 1e-6 : (d > 1e32 ? 1e32 : d); o->D2l[l] = d * d; }
      /* Cauchy point: alpha = |g~|^2 / (u^T H u), u = g / D^2 */
      gg = 0;
      for (int i = 0; i < n; i++) { o->tmp_c[i] = L->gc[i] / o->D2c[i]; gg += L->gc[i] * o->tmp_c[i]; }
      for (int l = 0; l < nl; l++) { o->tmp_l[l] = L->gl[l] / o->D2l[l]; gg += L->gl[l] * o->tmp_l[l]; }
      H_mul(o, L, o->tmp_c, o->tmp_l, o->step_c, o->step_l);
      double uHu = 0;
      for (int i = 0; i < n; i++) uHu += o->tmp_c[i] * o->step_c[i];
      for (int l = 0; l < nl; l++) uHu += o->tmp_l[l] * o->step_l[l];
      alpha = gg / uHu;
      int rc = solve_gn(o, L, mu, o->gn_c, o->gn_l, 0);
      if (rc) { /* linear solver failure -> invalid step */
        mu *= mu_inc; rep->total_iterations++;
        if (++invalid_run >= 5 || mu > max_mu * mu_inc) { rep->termination = D2BA_TERM_FAILURE; rep->succ = 0; break; }
        continue;
      }
      mu = fmax(min_mu, 2.0 * mu / mu_inc);
      nn = 0; gdn = 0;
      for (int i = 0; i < n; i++) { nn += o->gn_c[i] * o->gn_c[i] * o->D2c[i]; gdn += L->gc[i] * o->gn_c[i]; }
      for (int l = 0; l < nl; l++) { nn += o->gn_l[l] * o->gn_l[l] * o->D2l[l]; gdn += L->gl[l] * o->gn_l[l]; }
      reuse = 1;
    }
    /* traditional dogleg in the scaled space: step = c_g * (-g/D^2) ... expressed on unscaled vectors */
    double gn_norm = sqrt(nn), g_norm = sqrt(gg), c1, c2, step_norm;
    if (gn_norm <= radius) { c1 = 0; c2 = 1; step_norm = gn_norm; }
    else if (g_norm * alpha >= radius) { c1 = radius / g_norm; c2 = 0; step_norm = radius; }
    else {
      double b_dot_a = -alpha * gdn, a_sq = alpha * alpha * gg, b_minus_a_sq = nn - 2 * b_dot_a + a_sq;
      double c = b_dot_a - a_sq, d = sqrt(c * c + b_minus_a_sq * (radius * radius - a_sq));
      double beta = (c <= 0) ? (d - c) / b_minus_a_sq : (radius * radius - a_sq) / (d + c);
      c1 = alpha * (1 - beta); c2 = beta; step_norm = radius;
    }
    for (int i = 0; i < n; i++) o->step_c[i] = -c1 * L->gc[i] / o->D2c[i] + c2 * o->gn_c[i];
    for (int l = 0; l < nl; l++) o->step_l[l] = -c1 * L->gl[l] / o->D2l[l] + c2 * o->gn_l[l];
    /* model cost change = -(s^T g + 1/2 s^T H s) */
    H_mul(o, L, o->step_c, o->step_l, o->tmp_c, o->tmp_l);
    double sg = 0, sHs = 0;
    for (int i = 0; i < n; i++) { sg += o->step_c[i] * L->gc[i]; sHs += o->step_c[i] * o->tmp_c[i]; }
    for (int l = 0; l < nl; l++) { sg += o->step_l[l] * L->gl[l]; sHs += o->step_l[l] * o->tmp_l[l]; }
    double model_change = -(sg + 0.5 * sHs);
    rep->total_iterations++;
    if (!(model_change > 0)) { /* invalid step: StepIsInvalid -> mu *= 10, reuse = false */
      mu *= mu_inc; reuse = 0;
      if (++invalid_run >= 5) { rep->termination = D2BA_TERM_FAILURE; rep->succ = 0; break; }
      continue;
    }
    invalid_run = 0;
    apply_step(o, &o->x, o->step_c, o->step_l, &o->xc);
    double cand = linearize(o, &o->xc, L, 0);
    double x_norm, dx_norm;
    norms(o, &o->x, &o->xc, &x_norm, &dx_norm);
    if (!fixed) {
      if (dx_norm <= cf->parameter_tolerance * (x_norm + cf->parameter_tolerance)) { rep->termination = D2BA_TERM_PARAMETER_TOL; rep->total_iterations--; break; }
      if (fabs(cost - cand) <= cf->function_tolerance * cost) { rep->termination = D2BA_TERM_FUNCTION_TOL; rep->total_iterations--; break; }
    }
    double rel = (cost - cand) / model_change;
    if (rel > cf->min_relative_decrease) {
      copy_state(o, &o->x, &o->xc);
      cost = linearize(o, &o->x, L, 1);
      rep->successful_steps++;
      if (rel < 0.25) radius *= 0.5;
      if (rel > 0.75) radius = fmax(radius, 3.0 * step_norm);
      radius = fmin(cf->max_trust_region_radius, radius);
      reuse = 0;
      gmax = grad_max(o, L);
      if (!fixed && gmax <= cf->gradient_tolerance) { rep->termination = D2BA_TERM_GRADIENT_TOL; break; }
    } else {
      radius *= 0.5; reuse = 1;
      if (radius < 1e-32) { rep->termination = D2BA_TERM_FAILURE; rep->succ = 0; break; }
    }
  }
done:
  rep->final_cost = cost; rep->final_gradient_max_norm = gmax; rep->final_radius = radius;
  rep->total_time = now_s() - t0; rep->state_changes = 0;
}

static void ensure_cand(orc_handle *o) { (void)o; }
int orc_admm_solve(orc_handle **ag, int32_t n, int32_t fixed_mode, d2ba_report *reports);

int orc_solve(orc_handle *o, d2ba_report *rep) {
  d2ba_report r; memset(&r, 0, sizeof r);
  ensure_cand(o);
  if (o->admm_on) { orc_handle *one[1] = {o}; return orc_admm_solve(one, 1, 0, rep); }  /* a swarm of one */
  tr_solve(o, o->cfg.max_num_iterations, 0, &r);
  if (rep) *rep = r;
  return 0;
}
int orc_solve_fixed(orc_handle *o, int32_t iters, d2ba_report *rep) {
  d2ba_report r; memset(&r, 0, sizeof r);
  if (o->admm_on) {
    int keep = o->cfg.max_num_iterations; o->cfg.max_num_iterations = iters;
    orc_handle *one[1] = {o};
    int rc = orc_admm_solve(one, 1, 1, rep);
    o->cfg.max_num_iterations = keep;
    return rc;
  }
  tr_solve(o, iters, 1, &r);
  if (rep) *rep = r;
  return 0;
}

typedef struct { orc_handle **hs; int n, fixed_iters; d2ba_report *reports; volatile int *next; pthread_mutex_t *mu; } many_arg_t;
static void *many_worker(void *p) {
  many_arg_t *a = (many_arg_t *)p;
  for (;;) {
    pthread_mutex_lock(a->mu);
    int i = (*a->next)++;
    pthread_mutex_unlock(a->mu);
    if (i >= a->n) break;
    d2ba_report r; memset(&r, 0, sizeof r);
    if (a->fixed_iters > 0) tr_solve(a->hs[i], a->fixed_iters, 1, &r); else tr_solve(a->hs[i], a->hs[i]->cfg.max_num_iterations, 0, &r);
    if (a->reports) a->reports[i] = r;
  }
  return NULL;
}
int orc_solve_many(orc_handle **hs, int32_t n, int32_t nthreads, int32_t fixed_iters, d2ba_report *reports) {
  if (nthreads < 1) nthreads = 1;
  if (nthreads > 256) nthreads = 256;
  pthread_t th[256];
  pthread_mutex_t mu = PTHREAD_MUTEX_INITIALIZER;
  volatile int next = 0;
  many_arg_t a = {hs, n, fixed_iters, reports, &next, &mu};
  for (int t = 0; t < nthreads; t++) pthread_create(&th[t], NULL, many_worker, &a);
  for (int t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
  return 0;
}

/* ------------------------------------------------------------------ ADMM
 * ConsensusSolver::solve, d2common/src/solver/ConsensusSolver.cpp:39-75 with
 * syncData (:77-88), updateGlobal (:166-228), updateTilde (:108-164). */
int orc_admm_solve(orc_handle **ag, int32_t n, int32_t fixed_mode, d2ba_report *reports) {
  int max_steps = ag[0]->cfg.consensus_max_steps > 0 ? ag[0]->cfg.consensus_max_steps : 1;
  int n_slots = 0;
  for (int a = 0; a < n; a++) {
    orc_handle *o = ag[a];
    if (!o->cols_valid) assign_cols(o);
    if (o->n_slots > n_slots) n_slots = o->n_slots;
    /* ConsenusParamState::create: param_global = local value, tilde = 0 (ConsensusSolver.hpp:31-44);
     * solver->reset() before every solve clears them (ConsensusSolver.cpp:15-24) */
    memcpy(o->pose_z, o->x.pose, sizeof(double) * 7 * o->np); memcpy(o->ext_z, o->x.ext, sizeof(double) * 7 * o->ne);
    memset(o->pose_tilde, 0, sizeof(double) * 6 * o->np); memset(o->ext_tilde, 0, sizeof(double) * 6 * o->ne);
    if (reports) memset(reports + a, 0, sizeof(d2ba_report));
  }
  double *sum_p = (double *)malloc(sizeof(double) * 3 * (n_slots + 1)), *qs = (double *)malloc(sizeof(double) * 4 * n * (n_slots + 1));
  int *cnt = (int *)malloc(sizeof(int) * (n_slots + 1));
  double *zs = (double *)malloc(sizeof(double) * 7 * (n_slots + 1));
  for (int step = 0; step < max_steps; step++) {
    /* broadcastData + waitForSync + updateGlobal: z = average over agents holding the slot */
    memset(sum_p, 0, sizeof(double) * 3 * (n_slots + 1)); memset(cnt, 0, sizeof(int) * (n_slots + 1));
    for (int a = 0; a < n; a++) {
      orc_handle *o = ag[a];
      for (int pass = 0; pass < 2; pass++) {
        int nb = pass == 0 ? o->np : o->ne;
        for (int i = 0; i < nb; i++) {
          int s = pass == 0 ? o->pose_slot[i] : o->ext_slot[i];
          if (s < 0) continue;
          const double *xp = (pass == 0 ? o->x.pose : o->x.ext) + 7 * i;
          for (int q = 0; q < 3; q++) sum_p[3 * s + q] += xp[q];
          memcpy(qs + 4 * ((size_t)s * n + cnt[s]), xp + 3, 32);
          cnt[s]++;
        }
      }
    }
    for (int s = 0; s < n_slots; s++) {
      if (!cnt[s]) continue;
      /* Swarm::Pose::averagePoses (un-vendored swarm_msgs; ASSUMED = mean position + Utility::averageQuaterions) */
      for (int q = 0; q < 3; q++) zs[7 * s + q] = sum_p[3 * s + q] / cnt[s];
      orc_average_quats(cnt[s], qs + 4 * (size_t)s * n, zs + 7 * s + 3);
    }
    for (int a = 0; a < n; a++) {
      orc_handle *o = ag[a];
      for (int pass = 0; pass < 2; pass++) {
        int nb = pass == 0 ? o->np : o->ne;
        for (int i = 0; i < nb; i++) {
          int s = pass == 0 ? o->pose_slot[i] : o->ext_slot[i];
          if (s < 0) continue;
          double *z = (pass == 0 ? o->pose_z : o->ext_z) + 7 * i, *tl = (pass == 0 ? o->pose_tilde : o->ext_tilde) + 6 * i;
          const double *xp = (pass == 0 ? o->x.pose : o->x.ext) + 7 * i;
          memcpy(z, zs + 7 * s, 56);
          /* eigenvector sign is implementation-defined in the reference; fixed here to the
           * hemisphere of the local estimate (ASSUMED) so theta_err is the small rotation */
          double dot = z[3] * xp[3] + z[4] * xp[4] + z[5] * xp[5] + z[6] * xp[6];
          if (dot < 0) for (int q = 3; q < 7; q++) z[q] = -z[q];
          double d6[6];
          orc_delta_pose_tangent(z, xp, d6);
          for (int q = 0; q < 6; q++) tl[q] += (1.0 + o->cfg.relaxation_alpha) * d6[q];
        }
      }
      memcpy(o->lm_ref, o->x.lm, sizeof(double) * o->nl); memcpy(o->sb_ref, o->x.sb, sizeof(double) * 9 * o->nsb);
      o->td_ref = o->x.td;
    }
    /* solveLocalStep on every agent: max_num_iterations / max_steps (d2vins_params.cpp:156-158) */
    int iters = ag[0]->cfg.max_num_iterations / max_steps;
    if (iters < 1) iters = 1;
    for (int a = 0; a < n; a++) {
      d2ba_report r; memset(&r, 0, sizeof r);
      tr_solve(ag[a], iters, fixed_mode, &r);
      if (reports) {
        if (step == 0) reports[a].initial_cost = r.initial_cost;
        reports[a].total_iterations += r.total_iterations; reports[a].successful_steps += r.successful_steps;
        reports[a].final_cost = r.final_cost; reports[a].total_time += r.total_time; reports[a].succ = r.succ;
        reports[a].termination = r.termination; reports[a].final_gradient_max_norm = r.final_gradient_max_norm;
        reports[a].final_radius = r.final_radius;
      }
    }
  }
  free(sum_p); free(qs); free(cnt); free(zs);
  return 0;
}


/* many independent swarms on nthreads host threads (each swarm = n_agents handles solved by orc_admm_solve) */
typedef struct { orc_handle **hs; int n_swarms, n_agents, fixed; d2ba_report *reports; volatile int *next; pthread_mutex_t *mu; } swarm_arg_t;
static void *swarm_worker(void *p) {
  swarm_arg_t *a = (swarm_arg_t *)p;
  for (;;) {
    pthread_mutex_lock(a->mu);
    int i = (*a->next)++;
    pthread_mutex_unlock(a->mu);
    if (i >= a->n_swarms) break;
    orc_admm_solve(a->hs + (size_t)i * a->n_agents, a->n_agents, a->fixed, a->reports ? a->reports + (size_t)i * a->n_agents : NULL);
  }
  return NULL;
}
int orc_admm_many(orc_handle **hs, int32_t n_swarms, int32_t n_agents, int32_t nthreads, int32_t fixed_mode, d2ba_report *reports) {
  if (nthreads < 1) nthreads = 1;
  if (nthreads > 256) nthreads = 256;
  pthread_t th[256];
  pthread_mutex_t mu = PTHREAD_MUTEX_INITIALIZER;
  volatile int next = 0;
  swarm_arg_t a = {hs, n_swarms, n_agents, fixed_mode, reports, &next, &mu};
  for (int t = 0; t < nthreads; t++) pthread_create(&th[t], NULL, swarm_worker, &a);
  for (int t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
  return 0;
}

int orc_get_consensus(orc_handle *o, int32_t n, const d2ba_blockref *refs, double *z7, double *tilde6) {
  for (int i = 0; i < n; i++) {
    int k = find_block(o, refs[i].kind, refs[i].id);
    if (k < 0) return 1;
    if (refs[i].kind == D2BA_POSE) { memcpy(z7 + 7 * i, o->pose_z + 7 * k, 56); memcpy(tilde6 + 6 * i, o->pose_tilde + 6 * k, 48); }
    else { memcpy(z7 + 7 * i, o->ext_z + 7 * k, 56); memcpy(tilde6 + 6 * i, o->ext_tilde + 6 * k, 48); }
  }
  return 0;
}

int orc_get_blocks(orc_handle *o, int32_t kind, int32_t n, const int64_t *ids, double *out) {
  for (int i = 0; i < n; i++) {
    int k = find_block(o, kind, ids ? ids[i] : 0);
    if (k < 0) return 1;
    switch (kind) {
      case D2BA_POSE: memcpy(out + 7 * i, o->x.pose + 7 * k, 56); break;
      case D2BA_EXTRINSIC: memcpy(out + 7 * i, o->x.ext + 7 * k, 56); break;
      case D2BA_SPEED_BIAS: memcpy(out + 9 * i, o->x.sb + 9 * k, 72); break;
      case D2BA_TD: out[i] = o->x.td; break;
      case D2BA_LANDMARK: out[i] = o->x.lm[k]; break;
    }
  }
  return 0;
}

/* ------------------------------------------------------------------ debug */
int orc_debug_linearize(orc_handle *o) {
  if (!o->cols_valid) assign_cols(o);
  linearize(o, &o->x, &o->lin, 1);
  int n = o->n_c;
  for (int i = 0; i < n; i++) { double d = sqrt(o->lin.Hcc[(size_t)i * n + i]); d = d < 1e-6 ? 1e-6 : d; o->D2c[i] = d * d; }
  for (int l = 0; l < o->nl; l++) { double d = sqrt(o->lin.hl[l]); d = d < 1e-6 ? 1e-6 : d; o->D2l[l] = d * d; }
  solve_gn(o, &o->lin, 1e-8, o->gn_c, o->gn_l, 2);
  /* S now holds the un-factored reduced system; recompute the GN step separately */
  double *Skeep = (double *)malloc(sizeof(double) * (size_t)(n ? n : 1) * (n ? n : 1));
  memcpy(Skeep, o->S, sizeof(double) * (size_t)n * n);
  solve_gn(o, &o->lin, 1e-8, o->gn_c, o->gn_l, 0);
  memcpy(o->S, Skeep, sizeof(double) * (size_t)n * n);
  free(Skeep);
  return 0;
}

static int cmp_obs(const void *a, const void *b) {
  const obs_t *x = (const obs_t *)a, *y = (const obs_t *)b;
  int kx[5] = {x->type, x->pi, x->pj, x->ea, x->eb}, ky[5] = {y->type, y->pi, y->pj, y->ea, y->eb};
  for (int i = 0; i < 5; i++) if (kx[i] != ky[i]) return kx[i] < ky[i] ? -1 : 1;
  return x->seq < y->seq ? -1 : (x->seq > y->seq);
}

int orc_debug_get(orc_handle *o, int32_t item, void *out, int64_t out_bytes, int64_t *needed) {
  if (!o->cols_valid) assign_cols(o);
  int n = o->n_c, nlc = o->n_lc, nl = o->nl;
  const void *src = NULL; int64_t bytes = 0; int64_t tmp64; void *owned = NULL;
  switch (item) {
    case D2BA_DBG_N_CAM: tmp64 = n; src = &tmp64; bytes = 8; break;
    case D2BA_DBG_N_LC: tmp64 = nlc; src = &tmp64; bytes = 8; break;
    case D2BA_DBG_HCC: src = o->lin.Hcc; bytes = 8LL * n * n; break;
    case D2BA_DBG_GC: src = o->lin.gc; bytes = 8LL * n; break;
    case D2BA_DBG_HLL: src = o->lin.hl; bytes = 8LL * nl; break;
    case D2BA_DBG_GL: src = o->lin.gl; bytes = 8LL * nl; break;
    case D2BA_DBG_W: src = o->lin.W; bytes = 8LL * nl * nlc; break;
    case D2BA_DBG_COST: src = &o->lin.cost; bytes = 8; break;
    case D2BA_DBG_S: src = o->S; bytes = 8LL * n * n; break;
    case D2BA_DBG_GN_STEP: {
      double *v = (double *)malloc(8 * (size_t)(n + nl + 1));
      memcpy(v, o->gn_c, 8 * (size_t)n); memcpy(v + n, o->gn_l, 8 * (size_t)nl);
      owned = v; src = v; bytes = 8LL * (n + nl); break;
    }
    case D2BA_DBG_STEP: {
      double *v = (double *)malloc(8 * (size_t)(n + nl + 1));
      memcpy(v, o->step_c, 8 * (size_t)n); memcpy(v + n, o->step_l, 8 * (size_t)nl);
      owned = v; src = v; bytes = 8LL * (n + nl); break;
    }
    case D2BA_DBG_OBS_INDEX: {
      obs_t *s = (obs_t *)malloc(sizeof(obs_t) * (o->nobs + 1));
      memcpy(s, o->obs, sizeof(obs_t) * o->nobs);
      qsort(s, o->nobs, sizeof(obs_t), cmp_obs);
      int32_t *v = (int32_t *)malloc(24 * (size_t)(o->nobs + 1));
      for (int k = 0; k < o->nobs; k++) {
        v[6 * k] = s[k].type; v[6 * k + 1] = s[k].pi; v[6 * k + 2] = s[k].pj;
        v[6 * k + 3] = s[k].ea < 0 ? -1 : o->np + s[k].ea; v[6 * k + 4] = s[k].eb < 0 ? -1 : o->np + s[k].eb; v[6 * k + 5] = s[k].lm;
      }
      free(s); owned = v; src = v; bytes = 24LL * o->nobs; break;
    }
    case D2BA_DBG_COL_OF_BLOCK: {
      int cnt = o->np + o->ne + o->nsb + 1;
      int32_t *v = (int32_t *)malloc(4 * (size_t)cnt);
      int k = 0;
      for (int i = 0; i < o->np; i++) v[k++] = o->pose_col[i];
      for (int i = 0; i < o->ne; i++) v[k++] = o->ext_col[i];
      for (int i = 0; i < o->nsb; i++) v[k++] = o->sb_col[i];
      v[k++] = o->td_col;
      owned = v; src = v; bytes = 4LL * cnt; break;
    }
    case D2BA_DBG_PROJ_RESJAC: {
      /* raw (uncorrected) r[3] + J 3x26 per obs in input order */
      double *v = (double *)calloc((size_t)(o->nobs + 1) * 81, 8);
      double sq = o->cfg.focal_length / 1.5;
      for (int k = 0; k < o->nobs; k++) {
        const obs_t *t = o->obs + k; double *rec = v + (size_t)k * 81;
        double r[3] = {0}, Ji[21] = {0}, Jj[21] = {0}, Ja[21] = {0}, Jb[21] = {0}, Jl[3] = {0}, Jtd[3] = {0};
        int rows;
        if (t->type == ORC_PROJ_DEPTH_PRIOR) { rows = 1; orc_depth_prior_eval(o->x.lm[t->lm], t->depth, o->cfg.depth_sqrt_inf, r, Jl); }
        else {
          rows = t->type == ORC_PROJ_2F1C_DEPTH ? 3 : 2;
          orc_proj_eval(t->type, &t->c, sq, o->cfg.depth_sqrt_inf, t->pi >= 0 ? o->x.pose + 7 * t->pi : NULL,
                        t->pj >= 0 ? o->x.pose + 7 * t->pj : NULL, o->x.ext + 7 * t->ea, t->eb >= 0 ? o->x.ext + 7 * t->eb : NULL,
                        o->x.lm[t->lm], o->x.td, r, Ji, Jj, Ja, Jb, Jl, Jtd);
        }
        for (int q = 0; q < rows; q++) {
          rec[q] = r[q];
          double *Jr = rec + 3 + q * 26;
          for (int c = 0; c < 6; c++) { Jr[c] = Ji[q * 7 + c]; Jr[6 + c] = Jj[q * 7 + c]; Jr[12 + c] = Ja[q * 7 + c]; Jr[18 + c] = Jb[q * 7 + c]; }
          Jr[24] = Jl[q]; Jr[25] = Jtd[q];
        }
      }
      owned = v; src = v; bytes = 81LL * 8 * o->nobs; break;
    }
    default: return 1;
  }
  if (needed) *needed = bytes;
  int rc = 0;
  if (out) { if (out_bytes < bytes) rc = 2; else memcpy(out, src, (size_t)bytes); }
  free(owned);
  return rc;
}

/* ------------------------------------------------------------------ marginalization
 * Marginalizer::marginalize, d2vins/src/estimator/marginalization/marginalization.cpp:173-254, with
 * filterResiduals (:78-118), sortParams (:256-285), evaluate (:17-76) and the exact-inverse Schur complement of
 * Utility::schurComplement (d2common/include/d2common/utils.hpp:131-141, margin_sparse_solver: 1).
 * Configuration restated: remove_base_when_margin_remote = 2 (every landmark of a relevant residual is
 * marginalised, config/tum/tum_single.yaml:89), margin_enable_fej = 0 (:94) -> evaluation at the current state.
 * Kept blocks are ordered by ParamsType like sortParams (POSE, SPEED_BIAS, EXTRINSIC, TD); ties (std::sort over a
 * pointer-keyed map in the reference) are resolved by block insertion order here. Output: A dx = b information
 * form ("Ignore -b", :203-204), kept-block refs and their linearisation points x0. */
int orc_marginalize_x0(orc_handle *o, int32_t n_remove, const int64_t *remove_frame_ids, int32_t *m_out, int32_t max_m,
                       double *A_out, double *b_out, int32_t *nblk_out, int32_t max_blk, d2ba_blockref *refs_out,
                       double *x0_out) {
  if (!o->cols_valid) assign_cols(o);
  const state_t *x = &o->x;
  char *rem_pose = (char *)calloc(o->np + 1, 1), *rem_sb = (char *)calloc(o->nsb + 1, 1);
  for (int i = 0; i < o->np; i++) for (int k = 0; k < n_remove; k++) if (o->pose_id[i] == remove_frame_ids[k]) rem_pose[i] = 1;
  for (int i = 0; i < o->nsb; i++) for (int k = 0; k < n_remove; k++) if (o->sb_id[i] == remove_frame_ids[k]) rem_sb[i] = 1;
  /* relevant residuals (ParamResidualInfo.hpp relavant()) */
  char *rel_obs = (char *)calloc(o->nobs + 1, 1), *rel_imu = (char *)calloc(o->nimu + 1, 1);
  char *use_pose = (char *)calloc(o->np + 1, 1), *use_sb = (char *)calloc(o->nsb + 1, 1), *use_ext = (char *)calloc(o->ne + 1, 1),
       *use_lm = (char *)calloc(o->nl + 1, 1);
  int use_td = 0;
  for (int k = 0; k < o->nobs; k++) {
    const obs_t *t = o->obs + k;
    int r = 0;
    if (t->type == ORC_PROJ_DEPTH_PRIOR || t->type == ORC_PROJ_1F2C) {
      /* DepthResInfo / LandmarkOneFrameTwoCamResInfo::relavant: the anchor frame only (ParamResidualInfo.hpp:104-106,160-162) */
      r = t->fa >= 0 ? rem_pose[t->fa] : 0;
    } else r = rem_pose[t->pi] || rem_pose[t->pj];
    rel_obs[k] = (char)r;
    if (!r) continue;
    if (t->pi >= 0) { use_pose[t->pi] = 1; use_pose[t->pj] = 1; }
    if (t->type != ORC_PROJ_DEPTH_PRIOR) { use_ext[t->ea] = 1; if (t->eb >= 0) use_ext[t->eb] = 1; use_td = 1; }
    use_lm[t->lm] = 1;
  }
  for (int k = 0; k < o->nimu; k++) {
    const imu_t *t = o->imu + k;
    if (rem_pose[t->pi] || rem_pose[t->pj]) { rel_imu[k] = 1; use_pose[t->pi] = use_pose[t->pj] = 1; use_sb[t->si] = use_sb[t->sj] = 1; }
  }
  if (o->pm > 0)
    for (int i = 0; i < o->pnblk; i++) {
      int kind = o->pkind[i], idx = o->pindex[i];
      if (kind == D2BA_POSE) use_pose[idx] = 1; else if (kind == D2BA_EXTRINSIC) use_ext[idx] = 1; else if (kind == D2BA_SPEED_BIAS) use_sb[idx] = 1;
      else if (kind == D2BA_TD) use_td = 1; else use_lm[idx] = 1;
    }
  /* column layout: keep [POSE, SPEED_BIAS, EXTRINSIC, TD] | remove [POSE, SPEED_BIAS, LANDMARK] */
  int *cp = (int *)malloc(sizeof(int) * (o->np + 1)), *cs = (int *)malloc(sizeof(int) * (o->nsb + 1)), *ce = (int *)malloc(sizeof(int) * (o->ne + 1)),
      *cl = (int *)malloc(sizeof(int) * (o->nl + 1));
  int ctd = -1, c = 0, nblk = 0, xo = 0, rc = 0;
  for (int i = 0; i < o->np; i++) cp[i] = -1;
  for (int i = 0; i < o->nsb; i++) cs[i] = -1;
  for (int i = 0; i < o->ne; i++) ce[i] = -1;
  for (int i = 0; i < o->nl; i++) cl[i] = -1;
#define EMIT(kind_, id_, ptr_, sz_) do { if (nblk >= max_blk) { rc = 3; } else { refs_out[nblk].kind = kind_; refs_out[nblk].pad = 0; refs_out[nblk].id = id_; \
    if (x0_out) memcpy(x0_out + xo, ptr_, sizeof(double) * (sz_)); xo += (sz_); nblk++; } } while (0)
  for (int i = 0; i < o->np; i++) if (use_pose[i] && !rem_pose[i]) { cp[i] = c; c += 6; EMIT(D2BA_POSE, o->pose_id[i], x->pose + 7 * i, 7); }
  for (int i = 0; i < o->nsb; i++) if (use_sb[i] && !rem_sb[i]) { cs[i] = c; c += 9; EMIT(D2BA_SPEED_BIAS, o->sb_id[i], x->sb + 9 * i, 9); }
  for (int i = 0; i < o->ne; i++) if (use_ext[i]) { ce[i] = c; c += 6; EMIT(D2BA_EXTRINSIC, o->ext_id[i], x->ext + 7 * i, 7); }
  if (use_td) { ctd = c; c += 1; EMIT(D2BA_TD, 0, &x->td, 1); }
  const int nk = c;
  for (int i = 0; i < o->np; i++) if (use_pose[i] && rem_pose[i]) { cp[i] = c; c += 6; }
  for (int i = 0; i < o->nsb; i++) if (use_sb[i] && rem_sb[i]) { cs[i] = c; c += 9; }
  for (int i = 0; i < o->nl; i++) if (use_lm[i]) { cl[i] = c; c += 1; }
  const int N = c, nr = N - nk;
  if (rc == 0 && (nk > max_m)) rc = 4;
  if (rc == 0 && (nk == 0 || nr == 0)) rc = 5;   /* reference returns nullptr (:190-196) */
  if (rc) goto cleanup;
  {
    double *H = (double *)calloc((size_t)N * N, 8), *g = (double *)calloc(N, 8);
#define ACCJ(rows_, nb_, cols_, widths_, lds_, Js_, r_) do { \
      for (int a_ = 0; a_ < (nb_); a_++) { if ((cols_)[a_] < 0) continue; \
        for (int i_ = 0; i_ < (widths_)[a_]; i_++) { double gs_ = 0; for (int q_ = 0; q_ < (rows_); q_++) gs_ += (Js_)[a_][q_ * (lds_)[a_] + i_] * (r_)[q_]; g[(cols_)[a_] + i_] += gs_; } \
        for (int b_ = 0; b_ < (nb_); b_++) { if ((cols_)[b_] < 0) continue; \
          for (int i_ = 0; i_ < (widths_)[a_]; i_++) for (int j_ = 0; j_ < (widths_)[b_]; j_++) { double s_ = 0; \
            for (int q_ = 0; q_ < (rows_); q_++) s_ += (Js_)[a_][q_ * (lds_)[a_] + i_] * (Js_)[b_][q_ * (lds_)[b_] + j_]; \
            H[(size_t)((cols_)[a_] + i_) * N + (cols_)[b_] + j_] += s_; } } } } while (0)
    for (int k = 0; k < o->nobs; k++) {
      if (!rel_obs[k]) continue;
      const obs_t *t = o->obs + k;
      double r[3], Ji[21] = {0}, Jj[21] = {0}, Ja[21] = {0}, Jb[21] = {0}, Jl[3] = {0}, Jtd[3] = {0};
      int rows;
      eval_proj(o, x, t, 1, r, Ji, Jj, Ja, Jb, Jl, Jtd, &rows);
      int cols[6] = {t->pi >= 0 ? cp[t->pi] : -1, t->pj >= 0 ? cp[t->pj] : -1, t->ea >= 0 ? ce[t->ea] : -1, t->eb >= 0 ? ce[t->eb] : -1, cl[t->lm],
                     t->type == ORC_PROJ_DEPTH_PRIOR ? -1 : ctd};
      int widths[6] = {6, 6, 6, 6, 1, 1}, lds[6] = {7, 7, 7, 7, 1, 1};
      const double *Js[6] = {Ji, Jj, Ja, Jb, Jl, Jtd};
      ACCJ(rows, 6, cols, widths, lds, Js, r);
    }
    for (int k = 0; k < o->nimu; k++) {
      if (!rel_imu[k]) continue;
      const imu_t *t = o->imu + k;
      double r[15], Jpi[105], Jsi[135], Jpj[105], Jsj[135];
      orc_imu_eval(&t->c, o->cfg.gravity_norm, x->pose + 7 * t->pi, x->sb + 9 * t->si, x->pose + 7 * t->pj, x->sb + 9 * t->sj, r, Jpi, Jsi, Jpj, Jsj);
      int cols[4] = {cp[t->pi], cs[t->si], cp[t->pj], cs[t->sj]}, widths[4] = {6, 9, 6, 9}, lds[4] = {7, 9, 7, 9};
      const double *Js[4] = {Jpi, Jsi, Jpj, Jsj};
      ACCJ(15, 4, cols, widths, lds, Js, r);
    }
    if (o->pm > 0) {
      int m = o->pm;
      double *dx = (double *)malloc(8 * m), *r = (double *)malloc(8 * m);
      int *pc = (int *)malloc(sizeof(int) * o->pnblk), *pw = (int *)malloc(sizeof(int) * o->pnblk), *pl = (int *)malloc(sizeof(int) * o->pnblk);
      const double **pj = (const double **)malloc(sizeof(double *) * o->pnblk);
      for (int i = 0; i < o->pnblk; i++) {
        int kind = o->pkind[i], idx = o->pindex[i], off = o->poff[i];
        const double *x0 = o->px0 + 9 * i;
        if (kind == D2BA_POSE) { orc_prior_dx_pose(x->pose + 7 * idx, x0, dx + off); pc[i] = cp[idx]; }
        else if (kind == D2BA_EXTRINSIC) { orc_prior_dx_pose(x->ext + 7 * idx, x0, dx + off); pc[i] = ce[idx]; }
        else if (kind == D2BA_SPEED_BIAS) { for (int q = 0; q < 9; q++) dx[off + q] = x->sb[9 * idx + q] - x0[q]; pc[i] = cs[idx]; }
        else if (kind == D2BA_TD) { dx[off] = x->td - x0[0]; pc[i] = ctd; }
        else { dx[off] = x->lm[idx] - x0[0]; pc[i] = cl[idx]; }
        pw[i] = o->peff[i]; pl[i] = m; pj[i] = o->pJ + off;
      }
      for (int i = 0; i < m; i++) { double a = o->pe0[i]; for (int j = 0; j < m; j++) a += o->pJ[(size_t)i * m + j] * dx[j]; r[i] = a; }
      ACCJ(m, o->pnblk, pc, pw, pl, pj, r);
      free(dx); free(r); free(pc); free(pw); free(pl); free(pj);
    }
    /* A = H11 - H12 H22^-1 H21, b = g1 - H12 H22^-1 g2 with the exact inverse (Cholesky solve) */
    double *H22 = (double *)malloc(8 * (size_t)nr * nr), *X = (double *)malloc(8 * (size_t)nr * (nk + 1));
    for (int i = 0; i < nr; i++) for (int j = 0; j < nr; j++) H22[(size_t)i * nr + j] = H[(size_t)(nk + i) * N + nk + j];
    if (chol_inplace(H22, nr)) rc = 6;
    else {
      double *col = (double *)malloc(8 * nr);
      for (int j = 0; j <= nk; j++) {
        for (int i = 0; i < nr; i++) col[i] = j < nk ? H[(size_t)(nk + i) * N + j] : g[nk + i];
        chol_solve(H22, nr, col);
        for (int i = 0; i < nr; i++) X[(size_t)i * (nk + 1) + j] = col[i];
      }
      free(col);
      for (int i = 0; i < nk; i++) {
        for (int j = 0; j < nk; j++) {
          double s_ = H[(size_t)i * N + j];
          for (int k = 0; k < nr; k++) s_ -= H[(size_t)i * N + nk + k] * X[(size_t)k * (nk + 1) + j];
          A_out[(size_t)i * nk + j] = s_;
        }
        double s_ = g[i];
        for (int k = 0; k < nr; k++) s_ -= H[(size_t)i * N + nk + k] * X[(size_t)k * (nk + 1) + nk];
        b_out[i] = s_;
      }
      *m_out = nk; *nblk_out = nblk;
    }
    free(H22); free(X); free(H); free(g);
  }
cleanup:
  free(rem_pose); free(rem_sb); free(rel_obs); free(rel_imu); free(use_pose); free(use_sb); free(use_ext); free(use_lm);
  free(cp); free(cs); free(ce); free(cl);
  return rc;
}
