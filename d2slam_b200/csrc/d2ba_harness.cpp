// d2ba_harness.cpp -- C++ host-side stand-in for the caller of the hot path (D2VINS::D2Estimator).
//
// The reference rebuilds the solver problem for every solve: solver->reset(), setupImuFactors(),
// setupLandmarkFactors(), setupPriorFactor(), solver->solve(), state.syncFromState()
// (d2vins/src/estimator/d2estimator.cpp:604-685 and :502-602).  This harness keeps per-window host buffers
// (what D2EstimatorState and the factor objects hold) and replays exactly that sequence through the C ABI
// of libd2ba.so: d2ba_reset -> d2ba_set_blocks / d2ba_add_proj / d2ba_add_imu / d2ba_set_prior_info /
// d2ba_set_consensus (one host thread per group of windows, like one estimator thread per drone) ->
// d2ba_finalize (H2D) -> d2ba_solve_fixed -> d2ba_get_blocks (D2H).  bench.py uses it for the end-to-end
// number so that no Python overhead sits inside the timed region; it contains no numerics.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <deque>
#include <mutex>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#include "../../include/d2ba.h"

namespace {
struct HWin {
  std::vector<int64_t> frame_ids, cam_ids, sb_ids, lm_ids;
  std::vector<double> poses, ext, sb, inv_dep, prior_A, prior_b, prior_x0;
  std::vector<uint8_t> pose_const, ext_const;
  double td = 0; uint8_t td_const = 1;
  std::vector<d2ba_proj_obs> obs;
  std::vector<d2ba_imu> imu;
  std::vector<d2ba_blockref> prior_refs, cons_refs;
  std::vector<int32_t> slots; int n_slots = 0;
  // outputs
  std::vector<double> out_pose, out_sb, out_lm;
};
}  // namespace

struct rp_ctx { std::vector<HWin> win; double t_feed = 0, t_finalize = 0, t_solve = 0, t_fetch = 0; std::atomic<long long> ns_call[4]; };   // ns_call: thread-summed ns in set_blocks / add_proj / add_imu / prior+consensus

extern "C" {

rp_ctx *rp_create(int n_windows) { rp_ctx *c = new rp_ctx(); c->win.resize(n_windows); for (auto &a : c->ns_call) a = 0; return c; }
void rp_destroy(rp_ctx *c) { delete c; }

int rp_set_window(rp_ctx *c, int w, int np, const int64_t *frame_ids, const double *poses, const uint8_t *pose_const, int ne,
                  const int64_t *cam_ids, const double *ext, const uint8_t *ext_const, int nsb, const int64_t *sb_ids, const double *sb,
                  double td, int td_const, int nl, const int64_t *lm_ids, const double *inv_dep, int nobs, const d2ba_proj_obs *obs,
                  int nimu, const d2ba_imu *imu, int prior_m, const double *A, const double *b, int prior_nblk, const d2ba_blockref *refs,
                  int prior_x0_len, const double *x0, int ncons, const d2ba_blockref *crefs, const int32_t *slots, int n_slots) {
  if (!c || w < 0 || w >= (int)c->win.size()) return 1;
  HWin &W = c->win[w];
  W.frame_ids.assign(frame_ids, frame_ids + np); W.poses.assign(poses, poses + 7 * np); W.pose_const.assign(pose_const, pose_const + np);
  W.cam_ids.assign(cam_ids, cam_ids + ne); W.ext.assign(ext, ext + 7 * ne); W.ext_const.assign(ext_const, ext_const + ne);
  W.sb_ids.assign(sb_ids, sb_ids + nsb); W.sb.assign(sb, sb + 9 * nsb);
  W.td = td; W.td_const = (uint8_t)td_const;
  W.lm_ids.assign(lm_ids, lm_ids + nl); W.inv_dep.assign(inv_dep, inv_dep + nl);
  W.obs.assign(obs, obs + nobs); W.imu.assign(imu, imu + nimu);
  W.prior_A.clear(); W.prior_b.clear(); W.prior_refs.clear(); W.prior_x0.clear();
  if (prior_m > 0) { W.prior_A.assign(A, A + (size_t)prior_m * prior_m); W.prior_b.assign(b, b + prior_m); W.prior_refs.assign(refs, refs + prior_nblk); W.prior_x0.assign(x0, x0 + prior_x0_len); }
  W.cons_refs.clear(); W.slots.clear(); W.n_slots = n_slots;
  if (ncons > 0) { W.cons_refs.assign(crefs, crefs + ncons); W.slots.assign(slots, slots + ncons); }
  W.out_pose.assign(7 * np, 0.0); W.out_sb.assign(9 * nsb, 0.0); W.out_lm.assign(nl, 0.0);
  return 0;
}

static int feed_window(rp_ctx *c, d2ba_handle *h, int w, const HWin &W) {
  int rc;
  const int64_t zero = 0;
  auto t0 = std::chrono::steady_clock::now();
  auto lap = [&](int k) { auto t = std::chrono::steady_clock::now(); c->ns_call[k] += std::chrono::duration_cast<std::chrono::nanoseconds>(t - t0).count(); t0 = t; };
  if ((rc = d2ba_set_blocks(h, w, D2BA_POSE, (int)W.frame_ids.size(), W.frame_ids.data(), W.poses.data(), W.pose_const.data()))) return rc;
  if ((rc = d2ba_set_blocks(h, w, D2BA_EXTRINSIC, (int)W.cam_ids.size(), W.cam_ids.data(), W.ext.data(), W.ext_const.data()))) return rc;
  if ((rc = d2ba_set_blocks(h, w, D2BA_SPEED_BIAS, (int)W.sb_ids.size(), W.sb_ids.data(), W.sb.data(), nullptr))) return rc;
  if ((rc = d2ba_set_blocks(h, w, D2BA_TD, 1, &zero, &W.td, &W.td_const))) return rc;
  if ((rc = d2ba_set_blocks(h, w, D2BA_LANDMARK, (int)W.lm_ids.size(), W.lm_ids.data(), W.inv_dep.data(), nullptr))) return rc;
  lap(0);
  if ((rc = d2ba_add_proj(h, w, (int)W.obs.size(), W.obs.data()))) return rc;
  lap(1);
  if (!W.imu.empty() && (rc = d2ba_add_imu(h, w, (int)W.imu.size(), W.imu.data()))) return rc;
  lap(2);
  if (!W.prior_b.empty() && (rc = d2ba_set_prior_info(h, w, (int)W.prior_b.size(), W.prior_A.data(), W.prior_b.data(), (int)W.prior_refs.size(),
                                                     W.prior_refs.data(), W.prior_x0.data()))) return rc;
  if (!W.cons_refs.empty() && (rc = d2ba_set_consensus(h, w, (int)W.cons_refs.size(), W.cons_refs.data(), W.slots.data(), W.n_slots))) return rc;
  lap(3);
  return 0;
}

// Runs `steps` complete solve cycles; returns wall-clock seconds (steady_clock) or a negative error code.
double rp_run(rp_ctx *c, d2ba_handle *h, int steps, int iters, int nthreads, d2ba_report *last_reports) {
  if (!c || !h) return -1;
  const int nw = (int)c->win.size();
  if (nthreads < 1) nthreads = 1;
  if (nthreads > nw) nthreads = nw;
  std::vector<d2ba_report> reps(nw);
  auto t0 = std::chrono::steady_clock::now();
  c->t_feed = c->t_finalize = c->t_solve = c->t_fetch = 0;
  for (auto &a : c->ns_call) a = 0;
  auto now = []() { return std::chrono::steady_clock::now(); };
  auto secs = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) { return std::chrono::duration<double>(b - a).count(); };
  for (int s = 0; s < steps; s++) {
    auto ta = now();
    if (d2ba_reset(h)) return -2;
    std::atomic<int> next(0), err(0);
    auto work = [&]() { for (;;) { int w = next.fetch_add(1); if (w >= nw) break; int rc = feed_window(c, h, w, c->win[w]); if (rc) err = rc; } };
    if (nthreads == 1) work();
    else { std::vector<std::thread> th; for (int t = 0; t < nthreads; t++) th.emplace_back(work); for (auto &t : th) t.join(); }
    if (err) return -100 - err;
    auto tb = now();
    if (d2ba_finalize(h)) return -3;
    auto tc = now();
    if (d2ba_solve_fixed(h, iters, reps.data())) return -4;
    auto td = now();
    std::atomic<int> nx2(0);
    auto fetch = [&]() {
      for (;;) {
        int w = nx2.fetch_add(1); if (w >= nw) break;
        HWin &W = c->win[w];
        d2ba_get_blocks(h, w, D2BA_POSE, (int)W.frame_ids.size(), W.frame_ids.data(), W.out_pose.data());
        d2ba_get_blocks(h, w, D2BA_SPEED_BIAS, (int)W.sb_ids.size(), W.sb_ids.data(), W.out_sb.data());
        d2ba_get_blocks(h, w, D2BA_LANDMARK, (int)W.lm_ids.size(), W.lm_ids.data(), W.out_lm.data());
      }
    };
    if (nthreads == 1) fetch();
    else { std::vector<std::thread> th; for (int t = 0; t < nthreads; t++) th.emplace_back(fetch); for (auto &t : th) t.join(); }
    auto te = now();
    c->t_feed += secs(ta, tb); c->t_finalize += secs(tb, tc); c->t_solve += secs(tc, td); c->t_fetch += secs(td, te);
  }
  auto t1 = std::chrono::steady_clock::now();
  if (last_reports) memcpy(last_reports, reps.data(), sizeof(d2ba_report) * nw);
  return std::chrono::duration<double>(t1 - t0).count();
}

// Pipelined variant: the same per-step C-ABI sequence, but consecutive steps overlap like the stages of the reference's
// own estimator threads (front-end feeding frame k+1 while the back-end still solves frame k): stage 1 = reset + add,
// stage 2 = d2ba_finalize, stage 3 = d2ba_solve_fixed, stage 4 = d2ba_get_blocks, each step travelling through the stages on one
// of `n_handles` independent handles.  Every step still uploads all of its inputs and reads back all of its results.
namespace {
struct Chan {   // tiny blocking queue of step tokens
  std::mutex mu; std::condition_variable cv; std::deque<int> q;
  void push(int v) { { std::lock_guard<std::mutex> lk(mu); q.push_back(v); } cv.notify_one(); }
  int pop() { std::unique_lock<std::mutex> lk(mu); cv.wait(lk, [&] { return !q.empty(); }); int v = q.front(); q.pop_front(); return v; }
};
}  // namespace

double rp_run_pipelined(rp_ctx *c, d2ba_handle **hs, int n_handles, int steps, int iters, int nthreads, d2ba_report *last_reports) {
  if (!c || !hs || n_handles < 1) return -1;
  const int nw = (int)c->win.size();
  if (nthreads < 1) nthreads = 1;
  if (nthreads > nw) nthreads = nw;
  std::vector<d2ba_report> reps(nw);
  std::atomic<int> err(0);
  Chan free_h, to_finalize, to_solve;
  for (int i = 0; i < n_handles; i++) free_h.push(i);
  for (auto &a : c->ns_call) a = 0;
  auto t0 = std::chrono::steady_clock::now();
  auto now = []() { return std::chrono::steady_clock::now(); };
  auto secs = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) { return std::chrono::duration<double>(b - a).count(); };
  double busy_feed = 0, busy_fin = 0, busy_solve = 0, busy_fetch = 0;   // each written by exactly one stage thread
  std::thread th_fin([&]() {
    for (int s = 0; s < steps; s++) {
      int hi = to_finalize.pop();
      auto a = now();
      if (hi >= 0 && !err && d2ba_finalize(hs[hi])) err = 3;
      busy_fin += secs(a, now());
      to_solve.push(hi);
    }
  });
  Chan to_fetch;
  // two solves may be in flight (on different handles) so the GPU never idles across the host-side tail of a solve
  const char *ns_env = getenv("D2BA_RP_SOLVERS");
  const int n_solvers = ns_env ? std::max(1, atoi(ns_env)) : 1;
  std::mutex solve_mu;
  std::vector<std::thread> th_solve;
  for (int k = 0; k < n_solvers; k++)
    th_solve.emplace_back([&, k]() {
      std::vector<d2ba_report> my_reps(nw);
      for (;;) {
        int hi = to_solve.pop();
        if (hi < 0) break;   // poison pill
        auto a = now();
        if (!err && d2ba_solve_fixed(hs[hi], iters, my_reps.data())) err = 4;
        { std::lock_guard<std::mutex> lk(solve_mu); busy_solve += secs(a, now()); reps = my_reps; }
        to_fetch.push(hi);
      }
    });
  std::thread th_fetch([&]() {
    for (int s = 0; s < steps; s++) {
      int hi = to_fetch.pop();
      auto a = now();
      if (!err) {
        d2ba_handle *h = hs[hi];
        std::atomic<int> nx(0);
        auto fetch = [&]() {
          for (;;) {
            int w = nx.fetch_add(1); if (w >= nw) break;
            HWin &W = c->win[w];
            d2ba_get_blocks(h, w, D2BA_POSE, (int)W.frame_ids.size(), W.frame_ids.data(), W.out_pose.data());
            d2ba_get_blocks(h, w, D2BA_SPEED_BIAS, (int)W.sb_ids.size(), W.sb_ids.data(), W.out_sb.data());
            d2ba_get_blocks(h, w, D2BA_LANDMARK, (int)W.lm_ids.size(), W.lm_ids.data(), W.out_lm.data());
          }
        };
        std::vector<std::thread> th; for (int t = 0; t < std::min(nthreads, 4); t++) th.emplace_back(fetch); for (auto &t : th) t.join();
      }
      busy_fetch += secs(a, now());
      free_h.push(hi);
    }
  });
  for (int s = 0; s < steps; s++) {
    int hi = free_h.pop();
    d2ba_handle *h = hs[hi];
    auto a = now();
    if (!err) {
      if (d2ba_reset(h)) err = 2;
      std::atomic<int> next(0);
      auto work = [&]() { for (;;) { int w = next.fetch_add(1); if (w >= nw) break; int rc = feed_window(c, h, w, c->win[w]); if (rc) err = 100 + rc; } };
      std::vector<std::thread> th; for (int t = 0; t < nthreads; t++) th.emplace_back(work); for (auto &t : th) t.join();
    }
    busy_feed += secs(a, now());
    to_finalize.push(hi);
  }
  th_fin.join(); th_fetch.join();
  for (int k = 0; k < n_solvers; k++) to_solve.push(-1);
  for (auto &t : th_solve) t.join();
  c->t_feed = busy_feed; c->t_finalize = busy_fin; c->t_solve = busy_solve; c->t_fetch = busy_fetch;   // stage busy time (stages overlap)
  auto t1 = std::chrono::steady_clock::now();
  if (err) return -(double)err.load();
  if (last_reports) memcpy(last_reports, reps.data(), sizeof(d2ba_report) * nw);
  return std::chrono::duration<double>(t1 - t0).count();
}

void rp_breakdown(rp_ctx *c, double out[8]) { out[0] = c->t_feed; out[1] = c->t_finalize; out[2] = c->t_solve; out[3] = c->t_fetch; for (int k = 0; k < 4; k++) out[4 + k] = 1e-9 * (double)c->ns_call[k].load(); }

int rp_get_outputs(rp_ctx *c, int w, double *pose, double *sb, double *lm) {
  if (!c || w < 0 || w >= (int)c->win.size()) return 1;
  const HWin &W = c->win[w];
  if (pose) memcpy(pose, W.out_pose.data(), W.out_pose.size() * 8);
  if (sb) memcpy(sb, W.out_sb.data(), W.out_sb.size() * 8);
  if (lm) memcpy(lm, W.out_lm.data(), W.out_lm.size() * 8);
  return 0;
}

}  // extern "C"
