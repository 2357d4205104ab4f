// phylo_group: several GPUs behind one handle, for a single-process host such as the OCaml
// runtime (include/phylo_engine.h, last section). Built only on the public
// per-engine C ABI: one engine and one host worker thread per device, contiguous pattern shards
// on PHYLO_LNL_BLOCK boundaries, scalar results combined on the host (block partials folded by
// phylo_reduce_partials => lnL bit-identical to one engine; integer lengths summed exactly).
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <condition_variable>
#include <cstring>
#include <exception>
#include <functional>
#include <mutex>
#include <memory>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include "phylo_engine.h"

namespace {

// one persistent host thread per engine: CUDA calls of different devices overlap, and every
// engine is only ever touched by its own thread (the engine handle is not thread-safe)
struct Worker {
  std::thread th;
  std::mutex m;
  std::condition_variable cv;
  std::function<int()> job;
  bool has_job = false, done = false, quit = false;
  int rc = PHYLO_OK;
  std::string host_error;  // set when the job itself threw (the engine's message is not the cause then)

  Worker() {
    th = std::thread([this] {
      std::unique_lock<std::mutex> lk(m);
      for (;;) {
        cv.wait(lk, [this] { return has_job || quit; });
        if (quit) return;
        std::function<int()> j = std::move(job);
        has_job = false;
        lk.unlock();
        int r;
        std::string herr;
        try {
          r = j();
        } catch (const std::bad_alloc &) {  // report, never unwind out of the thread
          r = PHYLO_ERR_CUDA;
          herr = "host exception in worker: out of memory";
        } catch (const std::exception &ex) {
          r = PHYLO_ERR_CUDA;
          herr = std::string("host exception in worker: ") + ex.what();
        } catch (...) {
          r = PHYLO_ERR_CUDA;
          herr = "host exception in worker";
        }
        lk.lock();
        host_error = herr;
        rc = r;
        done = true;
        cv.notify_all();
      }
    });
  }
  void post(std::function<int()> j) {
    std::lock_guard<std::mutex> lk(m);
    job = std::move(j);
    has_job = true;
    done = false;
    cv.notify_all();
  }
  int wait() {
    std::unique_lock<std::mutex> lk(m);
    cv.wait(lk, [this] { return done; });
    return rc;
  }
  ~Worker() {
    {
      std::lock_guard<std::mutex> lk(m);
      quit = true;
      cv.notify_all();
    }
    if (th.joinable()) th.join();
  }
};

struct Shards {
  int64_t N = 0;
  std::vector<int64_t> lo, hi;
  bool loaded() const { return N > 0; }
};

// contiguous shards whose boundaries are multiples of `align` (the last one takes the ragged end)
static void cut(Shards &s, int64_t N, int n, int64_t align) {
  s.N = N;
  s.lo.assign(n, 0);
  s.hi.assign(n, 0);
  const int64_t blocks = (N + align - 1) / align, per = blocks / n, rem = blocks % n;
  for (int i = 0; i < n; ++i) {
    const int64_t lb = i * per + std::min<int64_t>(i, rem), hb = lb + per + (i < rem ? 1 : 0);
    s.lo[i] = std::min(lb * align, N);
    s.hi[i] = std::min(hb * align, N);
  }
}

}  // namespace

struct phylo_group {
  std::vector<phylo_engine *> eng;
  std::vector<Worker *> workers;
  std::string err;
  Shards lk, fitch;
  int S = 0, K = 0, felt = 1;
};

static std::string g_group_create_error;

static int gfail(phylo_group *g, int code, const std::string &msg) {
  if (g) g->err = msg; else g_group_create_error = msg;
  return code;
}

// run fn(i) on the worker of every shard for which active(i); first failure wins
static int fan_out(phylo_group *g, const std::function<bool(int)> &active, const std::function<int(int)> &fn,
                   const char *who) {
  const int n = (int)g->eng.size();
  std::vector<char> posted;
  try {  // nothing may unwind through the extern "C" entry points that call this
    posted.assign(n, 0);
    for (int i = 0; i < n; ++i)
      if (active(i)) {
        g->workers[i]->post([&fn, i] { return fn(i); });
        posted[i] = 1;
      }
  } catch (...) {
    for (int i = 0; i < (int)posted.size(); ++i)
      if (posted[i]) g->workers[i]->wait();
    return gfail(g, PHYLO_ERR_CUDA, std::string(who) + ": host exception while dispatching the shards (out of memory?)");
  }
  int first_rc = PHYLO_OK, first = -1;
  for (int i = 0; i < n; ++i)
    if (posted[i]) {
      const int rc = g->workers[i]->wait();
      if (rc != PHYLO_OK && first_rc == PHYLO_OK) { first_rc = rc; first = i; }
    }
  if (first_rc != PHYLO_OK) {
    const std::string &herr = g->workers[first]->host_error;
    return gfail(g, first_rc, std::string(who) + ": shard " + std::to_string(first) + ": " +
                                  (herr.empty() ? std::string(phylo_last_error(g->eng[first])) : herr));
  }
  return PHYLO_OK;
}

static bool all_shards(int) { return true; }

extern "C" int phylo_group_create(const int *devices, int n_devices, phylo_group **out) {
  if (!out) return gfail(nullptr, PHYLO_ERR_ARG, "phylo_group_create: out is NULL");
  *out = nullptr;
  if (!devices || n_devices < 1 || n_devices > 64)
    return gfail(nullptr, PHYLO_ERR_ARG, "phylo_group_create: need 1..64 devices");
  phylo_group *g = new phylo_group();
  for (int i = 0; i < n_devices; ++i) {
    phylo_engine *e = nullptr;
    const int rc = phylo_engine_create(devices[i], &e);
    if (rc != PHYLO_OK) {
      const std::string msg = std::string("phylo_group_create: device ") + std::to_string(devices[i]) + ": " + phylo_last_error(nullptr);
      phylo_group_destroy(g);
      return gfail(nullptr, rc, msg);
    }
    g->eng.push_back(e);
    g->workers.push_back(new Worker());
  }
  *out = g;
  return PHYLO_OK;
}

extern "C" void phylo_group_destroy(phylo_group *g) {
  if (!g) return;
  for (Worker *w : g->workers) delete w;  // joins: no call is in flight when destroy is legal
  for (phylo_engine *e : g->eng) phylo_engine_destroy(e);
  delete g;
}

extern "C" const char *phylo_group_last_error(const phylo_group *g) {
  return g ? g->err.c_str() : g_group_create_error.c_str();
}

extern "C" int phylo_group_size(const phylo_group *g) { return g ? (int)g->eng.size() : 0; }

extern "C" phylo_engine *phylo_group_engine(phylo_group *g, int i) {
  return (g && i >= 0 && i < (int)g->eng.size()) ? g->eng[i] : nullptr;
}

extern "C" int phylo_group_shard(const phylo_group *g, int which, int i, int64_t *lo, int64_t *hi) {
  if (!g || i < 0 || i >= (int)g->eng.size() || (which != 0 && which != 1)) return PHYLO_ERR_ARG;
  const Shards &s = which ? g->fitch : g->lk;
  if (!s.loaded()) return PHYLO_ERR_STATE;
  if (lo) *lo = s.lo[i];
  if (hi) *hi = s.hi[i];
  return PHYLO_OK;
}

extern "C" int phylo_group_set_option(phylo_group *g, int option, int64_t value) {
  if (!g) return PHYLO_ERR_ARG;
  for (size_t i = 0; i < g->eng.size(); ++i) {
    const int rc = phylo_engine_set_option(g->eng[i], option, value);
    if (rc != PHYLO_OK) return gfail(g, rc, std::string("group_set_option: ") + phylo_last_error(g->eng[i]));
  }
  return PHYLO_OK;
}

extern "C" int phylo_group_set_symbol_table(phylo_group *g, const uint64_t *table256) {
  if (!g) return PHYLO_ERR_ARG;
  for (size_t i = 0; i < g->eng.size(); ++i) {
    const int rc = phylo_engine_set_symbol_table(g->eng[i], table256);
    if (rc != PHYLO_OK) return gfail(g, rc, std::string("group_set_symbol_table: ") + phylo_last_error(g->eng[i]));
  }
  return PHYLO_OK;
}

// ------------------------------------------------------------------- likelihood ----
extern "C" int phylo_group_lk_set_model(phylo_group *g, int S, int K, const double *U, const double *D,
                                        const double *Ui, const double *priors, const double *rates,
                                        const double *probs, double pinvar) {
  if (!g) return PHYLO_ERR_ARG;
  const int rc = fan_out(g, all_shards, [&](int i) {
    return phylo_lk_set_model(g->eng[i], S, K, U, D, Ui, priors, rates, probs, pinvar);
  }, "group_lk_set_model");
  if (rc == PHYLO_OK) { g->S = S; g->K = K; }
  return rc;
}

extern "C" int phylo_group_lk_set_tips(phylo_group *g, int T, int64_t N, const void *masks, int mask_bytes,
                                       const double *weights, int capacity) {
  if (!g) return PHYLO_ERR_ARG;
  if (N < 1 || !masks || mask_bytes < 1) return gfail(g, PHYLO_ERR_ARG, "group_lk_set_tips: bad arguments");
  g->lk = Shards();
  Shards s;
  cut(s, N, (int)g->eng.size(), PHYLO_LNL_BLOCK);
  const int rc = fan_out(g, [&](int i) { return s.hi[i] > s.lo[i]; }, [&](int i) {
    return phylo_lk_set_tips_pitched(g->eng[i], T, s.hi[i] - s.lo[i], (const char *)masks + (size_t)s.lo[i] * mask_bytes,
                                     mask_bytes, (uint64_t)N * mask_bytes, weights ? weights + s.lo[i] : nullptr, capacity);
  }, "group_lk_set_tips");
  if (rc == PHYLO_OK) g->lk = s;
  return rc;
}

static bool lk_active(const phylo_group *g, int i) { return g->lk.hi[i] > g->lk.lo[i]; }

extern "C" int phylo_group_lk_score_tree(phylo_group *g, const phylo_op *ops, int n_ops, int root_a, int root_b,
                                         double root_t, double *lnl_out) {
  if (!g || !lnl_out) return PHYLO_ERR_ARG;
  if (!g->lk.loaded()) return gfail(g, PHYLO_ERR_STATE, "group_lk_score_tree: no tips loaded");
  const int n = (int)g->eng.size();
  // level-1 block partials of every shard land in their place of the whole alignment's list
  const int64_t n_blocks = (g->lk.N + PHYLO_LNL_BLOCK - 1) / PHYLO_LNL_BLOCK;
  std::vector<double> partials((size_t)n_blocks, 0.0);
  std::vector<double> shard_lnl(n, 0.0);
  const int rc = fan_out(g, [&](int i) { return lk_active(g, i); }, [&](int i) {
    int r = phylo_lk_score_tree(g->eng[i], ops, n_ops, root_a, root_b, root_t, &shard_lnl[i]);
    if (r != PHYLO_OK) return r;
    int64_t nb = 0;
    return phylo_lk_get_block_partials(g->eng[i], partials.data() + g->lk.lo[i] / PHYLO_LNL_BLOCK, &nb);
  }, "group_lk_score_tree");
  if (rc != PHYLO_OK) return rc;
  *lnl_out = phylo_reduce_partials(partials.data(), n_blocks);
  return PHYLO_OK;
}

extern "C" int phylo_group_lk_edge_lnl(phylo_group *g, int a, int b, const double *t, int n_t, double *lnl_out) {
  if (!g || !t || n_t < 1 || !lnl_out) return PHYLO_ERR_ARG;
  if (!g->lk.loaded()) return gfail(g, PHYLO_ERR_STATE, "group_lk_edge_lnl: no tips loaded");
  const int n = (int)g->eng.size();
  std::vector<double> vals((size_t)n * n_t, 0.0);
  const int rc = fan_out(g, [&](int i) { return lk_active(g, i); }, [&](int i) {
    return phylo_lk_edge_lnl(g->eng[i], a, b, t, n_t, vals.data() + (size_t)i * n_t);
  }, "group_lk_edge_lnl");
  if (rc != PHYLO_OK) return rc;
  for (int j = 0; j < n_t; ++j) {
    double s = 0.0;
    for (int i = 0; i < n; ++i) s += vals[(size_t)i * n_t + j];
    lnl_out[j] = s;
  }
  return PHYLO_OK;
}

// lnL, d1, d2 at one length, summed over the shards in shard order
static int group_edge_eval(phylo_group *g, double t, double *lnl, double *d1, double *d2) {
  const int n = (int)g->eng.size();
  std::vector<double> v((size_t)3 * n, 0.0);
  const int rc = fan_out(g, [&](int i) { return lk_active(g, i); }, [&](int i) {
    return phylo_lk_edge_eval(g->eng[i], &t, 1, &v[3 * i], &v[3 * i + 1], &v[3 * i + 2]);
  }, "group_lk_optimize_branch");
  if (rc != PHYLO_OK) return rc;
  *lnl = *d1 = *d2 = 0.0;
  for (int i = 0; i < n; ++i) { *lnl += v[3 * i]; *d1 += v[3 * i + 1]; *d2 += v[3 * i + 2]; }
  return PHYLO_OK;
}

// Same driver as phylo_lk_optimize_branch (engine.cu), on the sums: the bracket shrinks with the
// sign of dlnL/dt; a Newton step that leaves it falls back to regula falsi on the derivative
// (kept off the ends) or, before a sign change has been seen, to the geometric mean.
extern "C" int phylo_group_lk_optimize_branch(phylo_group *g, int a, int b, double t0, double t_min, double t_max,
                                              double tol, int max_iter, double *t_opt, double *lnl_opt,
                                              int *iters_out) {
  if (!g) return PHYLO_ERR_ARG;
  if (!g->lk.loaded()) return gfail(g, PHYLO_ERR_STATE, "group_lk_optimize_branch: no tips loaded");
  if (!(t_min > 0.0) || !(t_max > t_min) || !t_opt)
    return gfail(g, PHYLO_ERR_ARG, "group_lk_optimize_branch: need 0 < t_min < t_max");
  int rc = fan_out(g, [&](int i) { return lk_active(g, i); }, [&](int i) {
    return phylo_lk_edge_prepare(g->eng[i], a, b);
  }, "group_lk_optimize_branch");
  if (rc != PHYLO_OK) return rc;
  double lo = t_min, hi = t_max, t = std::min(std::max(t0, t_min), t_max);
  double glo = 0.0, ghi = 0.0;
  bool has_lo = false, has_hi = false;
  double lnl = 0.0, d1 = 0.0, d2 = 0.0, best_t = t, best_lnl = -INFINITY;
  int it = 0;
  if (tol <= 0.0) tol = 1e-8;
  if (max_iter < 1) max_iter = 50;
  for (; it < max_iter; ++it) {
    if ((rc = group_edge_eval(g, t, &lnl, &d1, &d2)) != PHYLO_OK) return rc;
    if (lnl > best_lnl) { best_lnl = lnl; best_t = t; }
    if (d1 > 0.0) { lo = t; glo = d1; has_lo = true; } else { hi = t; ghi = d1; has_hi = true; }
    double next = (d2 < 0.0) ? t - d1 / d2 : -1.0;
    if (!(next > lo && next < hi)) {
      if (has_lo && has_hi) {
        next = lo - glo * (hi - lo) / (ghi - glo);
        const double margin = 0.05 * (hi - lo);
        next = std::min(std::max(next, lo + margin), hi - margin);
      } else {
        next = std::sqrt(lo * hi);
      }
    }
    const bool at_bound = (t <= t_min && d1 <= 0.0) || (t >= t_max && d1 >= 0.0);
    if (at_bound || std::fabs(next - t) <= tol * std::max(t, 1e-8) || hi - lo <= tol * std::max(lo, 1e-8)) { ++it; break; }
    t = next;
  }
  *t_opt = best_t;
  if (lnl_opt) *lnl_opt = best_lnl;
  if (iters_out) *iters_out = it;
  return PHYLO_OK;
}

extern "C" int phylo_group_lk_get_site_lnl(phylo_group *g, double *out) {
  if (!g || !out) return PHYLO_ERR_ARG;
  if (!g->lk.loaded()) return gfail(g, PHYLO_ERR_STATE, "group_lk_get_site_lnl: no tips loaded");
  return fan_out(g, [&](int i) { return lk_active(g, i); }, [&](int i) {
    return phylo_lk_get_site_lnl(g->eng[i], out + g->lk.lo[i]);
  }, "group_lk_get_site_lnl");
}

extern "C" int phylo_group_lk_get_clv(phylo_group *g, int node, double *clv_out, int32_t *scale_out) {
  if (!g || !clv_out) return PHYLO_ERR_ARG;
  if (!g->lk.loaded()) return gfail(g, PHYLO_ERR_STATE, "group_lk_get_clv: no tips loaded");
  const size_t ks = (size_t)g->K * g->S;
  return fan_out(g, [&](int i) { return lk_active(g, i); }, [&](int i) {
    return phylo_lk_get_clv(g->eng[i], node, clv_out + (size_t)g->lk.lo[i] * ks,
                            scale_out ? scale_out + g->lk.lo[i] : nullptr);
  }, "group_lk_get_clv");
}

// ------------------------------------------------------------------------ Fitch ----
static bool fitch_active(const phylo_group *g, int i) { return g->fitch.hi[i] > g->fitch.lo[i]; }

extern "C" int phylo_group_fitch_set_tips(phylo_group *g, int T, int64_t N, int elt_bytes, int n_states,
                                          const void *codes, const double *weights, int capacity) {
  if (!g) return PHYLO_ERR_ARG;
  if (N < 1 || !codes || elt_bytes < 1) return gfail(g, PHYLO_ERR_ARG, "group_fitch_set_tips: bad arguments");
  g->fitch = Shards();
  Shards s;
  cut(s, N, (int)g->eng.size(), PHYLO_LNL_BLOCK);  // multiples of 1024 characters = whole 32-word tiles
  const int rc = fan_out(g, [&](int i) { return s.hi[i] > s.lo[i]; }, [&](int i) {
    return phylo_fitch_set_tips_pitched(g->eng[i], T, s.hi[i] - s.lo[i], elt_bytes, n_states,
                                        (const char *)codes + (size_t)s.lo[i] * elt_bytes, (uint64_t)N * elt_bytes,
                                        weights ? weights + s.lo[i] : nullptr, capacity);
  }, "group_fitch_set_tips");
  if (rc == PHYLO_OK) { g->fitch = s; g->felt = elt_bytes; }
  return rc;
}

extern "C" int phylo_group_fitch_score_tree(phylo_group *g, const phylo_op *ops, int n_ops, int root_a,
                                            int root_b, uint64_t *length_out) {
  if (!g || !length_out) return PHYLO_ERR_ARG;
  if (!g->fitch.loaded()) return gfail(g, PHYLO_ERR_STATE, "group_fitch_score_tree: no Fitch data loaded");
  const int n = (int)g->eng.size();
  std::vector<uint64_t> len(n, 0);
  const int rc = fan_out(g, [&](int i) { return fitch_active(g, i); }, [&](int i) {
    return phylo_fitch_score_tree(g->eng[i], ops, n_ops, root_a, root_b, &len[i]);
  }, "group_fitch_score_tree");
  if (rc != PHYLO_OK) return rc;
  uint64_t total = 0;
  for (int i = 0; i < n; ++i) total += len[i];
  *length_out = total;
  return PHYLO_OK;
}

extern "C" int phylo_group_fitch_uppass(phylo_group *g, const phylo_op *ops, int n_ops, int root_a, int root_b) {
  if (!g) return PHYLO_ERR_ARG;
  if (!g->fitch.loaded()) return gfail(g, PHYLO_ERR_STATE, "group_fitch_uppass: no Fitch data loaded");
  return fan_out(g, [&](int i) { return fitch_active(g, i); }, [&](int i) {
    return phylo_fitch_uppass(g->eng[i], ops, n_ops, root_a, root_b);
  }, "group_fitch_uppass");
}

extern "C" int phylo_group_fitch_get_states(phylo_group *g, int node, int which, void *out) {
  if (!g || !out) return PHYLO_ERR_ARG;
  if (!g->fitch.loaded()) return gfail(g, PHYLO_ERR_STATE, "group_fitch_get_states: no Fitch data loaded");
  return fan_out(g, [&](int i) { return fitch_active(g, i); }, [&](int i) {
    return phylo_fitch_get_states(g->eng[i], node, which, (char *)out + (size_t)g->fitch.lo[i] * g->felt);
  }, "group_fitch_get_states");
}

// Site-pattern compression over all devices: every device compresses a contiguous slab of
// sites (first stage, concurrently), then the per-slab pattern tables -- tiny next to the alignment -- are merged by
// one more pass on device 0 with the slabs' weights as input weights (second stage). Slabs are in site order and each
// table is in first-occurrence order, so the merged table is in the alignment's first-occurrence order: patterns,
// weights (exact for the integer-valued weights the reference uses) and the site map are identical to what one
// device returns for the whole alignment.
extern "C" int phylo_group_compress_patterns(phylo_group *g, int T, int64_t N, const void *masks, int mask_bytes,
                                             const double *weights_in, void *patterns_out, double *weights_out,
                                             int32_t *site_to_pattern, int64_t *n_patterns) {
  if (!g) return PHYLO_ERR_ARG;
  if (T < 1 || N < 1 || !masks || !patterns_out || !weights_out || !n_patterns ||
      !(mask_bytes == 1 || mask_bytes == 2 || mask_bytes == 4 || mask_bytes == 8))
    return gfail(g, PHYLO_ERR_ARG, "group_compress_patterns: bad arguments");
  const int n = (int)g->eng.size();
  if (n == 1 || N < 4096 * (int64_t)n)
    return phylo_compress_patterns(g->eng[0], T, N, masks, mask_bytes, weights_in, patterns_out, weights_out,
                                   site_to_pattern, n_patterns) == PHYLO_OK
               ? PHYLO_OK
               : gfail(g, PHYLO_ERR_CUDA, std::string("group_compress_patterns: ") + phylo_last_error(g->eng[0]));
  try {
    static const bool timing = [] { const char *v = getenv("PHYLO_GROUP_TIMING"); return v && v[0] == '1'; }();
    const auto tt0 = std::chrono::steady_clock::now();
    auto lap = [&](const char *what) {
      if (timing) fprintf(stderr, "[group compress] %s: +%.1f ms\n", what, std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - tt0).count());
    };
    Shards sh;
    cut(sh, N, n, 1024);
    const size_t EB = (size_t)mask_bytes, pitch = (size_t)N * EB;
    // worst-case sized but never zero-filled: only the P_i patterns a slab really has are ever touched
    std::vector<std::unique_ptr<unsigned char[]>> pats(n);
    std::vector<std::unique_ptr<double[]>> wts(n);
    std::vector<std::unique_ptr<int32_t[]>> maps(n);
    std::vector<int64_t> np(n, 0);
    int rc = fan_out(g, [&](int i) { return sh.hi[i] > sh.lo[i]; }, [&](int i) {
      const int64_t ns = sh.hi[i] - sh.lo[i];
      pats[i].reset(new unsigned char[(size_t)T * ns * EB]);
      wts[i].reset(new double[(size_t)ns]);
      maps[i].reset(new int32_t[(size_t)ns]);
      return phylo_compress_patterns_pitched(g->eng[i], T, ns, (const char *)masks + (size_t)sh.lo[i] * EB, mask_bytes, pitch,
                                             weights_in ? weights_in + sh.lo[i] : nullptr, pats[i].get(), wts[i].get(),
                                             maps[i].get(), &np[i]);
    }, "group_compress_patterns");
    if (rc != PHYLO_OK) return rc;
    lap("stage 1 done");
    // second stage: the slabs' tables side by side ([T][sum of P_i]), their weights as input weights
    int64_t tot = 0;
    std::vector<int64_t> off(n + 1, 0);
    for (int i = 0; i < n; ++i) { off[i] = tot; tot += np[i]; }
    off[n] = tot;
    std::vector<unsigned char> cat((size_t)T * tot * EB);
    std::vector<double> wcat((size_t)tot);
    for (int i = 0; i < n; ++i) {
      for (int t = 0; t < T; ++t)
        std::memcpy(cat.data() + ((size_t)t * tot + off[i]) * EB, pats[i].get() + (size_t)t * np[i] * EB, (size_t)np[i] * EB);
      std::copy(wts[i].get(), wts[i].get() + np[i], wcat.begin() + off[i]);
    }
    lap("tables concatenated");
    std::vector<int32_t> map2((size_t)tot);
    std::vector<unsigned char> out2((size_t)T * tot * EB);
    std::vector<double> w2((size_t)tot);
    int64_t P = 0;
    rc = fan_out(g, [&](int i) { return i == 0; }, [&](int) {
      return phylo_compress_patterns(g->eng[0], T, tot, cat.data(), mask_bytes, wcat.data(), out2.data(), w2.data(), map2.data(), &P);
    }, "group_compress_patterns");
    if (rc != PHYLO_OK) return rc;
    lap("stage 2 done");
    std::memcpy(patterns_out, out2.data(), (size_t)T * P * EB);
    std::copy(w2.begin(), w2.begin() + P, weights_out);
    if (site_to_pattern)
      for (int i = 0; i < n; ++i)
        for (int64_t s = sh.lo[i]; s < sh.hi[i]; ++s) site_to_pattern[s] = map2[(size_t)(off[i] + maps[i][(size_t)(s - sh.lo[i])])];
    *n_patterns = P;
    lap("site map merged");
    return PHYLO_OK;
  } catch (const std::exception &ex) {
    return gfail(g, PHYLO_ERR_CUDA, std::string("group_compress_patterns: host exception: ") + ex.what());
  }
}

