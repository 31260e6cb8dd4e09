// TEST INFRASTRUCTURE ONLY -- never linked into, imported by, or called from the product path.
//
// oracle/_ref/libref_oracle.so: the UNMODIFIED reference hot path (headers under
// /root/reference/include, compiled where they lie) behind a tiny C API so that pytest and the
// bench's CPU-baseline leg can drive it through ctypes. Nothing from the reference is copied; this
// file only *calls* reference classes. The glue below follows python/include/index.hpp (the
// pybind layer, which cannot be compiled here because it drags in Eigen/OpenMP through RaBitQ):
//   fit           -> index.hpp:177-227
//   search        -> index.hpp:236-258
//   batch_search  -> index.hpp:289-346  (Scheduler + coroutine search + rerank)
//   rerank        -> index.hpp:450-488
//   save / load   -> index.hpp:113-175
//   insert/remove -> index.hpp:229-234
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <memory>
#include <queue>
#include <string>
#include <thread>
#include <vector>

#include "executor/jobs/graph_search_job.hpp"
#include "executor/jobs/graph_update_job.hpp"
#include "executor/scheduler.hpp"
#include "index/graph/hnsw/hnsw_builder.hpp"
#include "index/graph/knng/nndescent.hpp"
#include "index/graph/nsg/nsg_builder.hpp"
#include "simd/cpu_features.hpp"
#include "simd/distance_ip.hpp"
#include "simd/distance_l2.hpp"
#include "space/raw_space.hpp"
#include "space/sq4_space.hpp"
#include "space/sq8_space.hpp"
#include "utils/evaluate.hpp"

using namespace alaya;  // NOLINT

namespace {

thread_local std::string g_err;

struct RefIndexBase {
  virtual ~RefIndexBase() = default;
  virtual void fit(float *data, uint32_t n, uint32_t efc, uint32_t threads) = 0;
  virtual void fit_nsg(float *data, uint32_t n, uint32_t efc) = 0;
  virtual void knng(float *data, uint32_t n, uint32_t k, uint32_t *out) = 0;
  virtual void search(float *q, uint32_t topk, uint32_t ef, uint32_t *ids) = 0;
  virtual void search_updated(float *q, uint32_t topk, uint32_t ef, uint32_t *ids) = 0;
  virtual void batch_search(float *q, uint32_t nq, uint32_t topk, uint32_t ef, uint32_t threads,
                            uint32_t *ids) = 0;
  virtual void batch_search_solo(float *q, uint32_t nq, uint32_t topk, uint32_t ef,
                                 uint32_t threads, uint32_t *ids) = 0;
  virtual void batch_search_with_distance(float *q, uint32_t nq, uint32_t topk, uint32_t ef,
                                          uint32_t threads, uint32_t *ids, float *dist,
                                          int *has_dist) = 0;
  virtual void traverse(float *q, uint32_t ef, uint32_t *pool_ids, float *pool_dist,
                        uint32_t *pool_size, uint64_t *stats) = 0;
  virtual uint32_t insert(float *v, uint32_t ef) = 0;
  virtual void remove(uint32_t id) = 0;
  virtual void save(const char *ip, const char *dp, const char *qp) = 0;
  virtual void load(const char *ip, const char *dp, const char *qp) = 0;
  virtual Graph<float, uint32_t> *graph() = 0;
  virtual RawSpace<> *raw() = 0;
  virtual const uint8_t *code_row(uint32_t id) = 0;
  virtual int quant_minmax(float *mn, float *mx) = 0;
  virtual float search_space_distance(uint32_t i, uint32_t j) = 0;
  uint32_t dim_{0};
};

template <typename SearchSpace>
struct RefIndex : RefIndexBase {
  using BuildSpace = RawSpace<>;
  static constexpr bool kSame = std::is_same_v<SearchSpace, BuildSpace>;

  MetricType metric_;
  uint32_t capacity_;
  uint32_t max_nbrs_{32};  // params_.max_nbrs_ is never set from Python (params.hpp:36)
  std::shared_ptr<Graph<float, uint32_t>> graph_;
  std::shared_ptr<BuildSpace> build_space_;
  std::shared_ptr<SearchSpace> search_space_;
  std::shared_ptr<GraphSearchJob<SearchSpace>> search_job_;
  std::shared_ptr<GraphUpdateJob<SearchSpace>> update_job_;
  std::shared_ptr<JobContext<uint32_t>> job_context_;

  RefIndex(MetricType m, uint32_t cap) : metric_(m), capacity_(cap) {}

  void make_jobs() {
    job_context_ = std::make_shared<JobContext<uint32_t>>();
    search_job_ = std::make_shared<GraphSearchJob<SearchSpace>>(search_space_, graph_, job_context_);
    update_job_ = std::make_shared<GraphUpdateJob<SearchSpace>>(search_job_);
  }

  void fit(float *data, uint32_t n, uint32_t efc, uint32_t threads) override {
    build_space_ = std::make_shared<BuildSpace>(capacity_, dim_, metric_);
    build_space_->fit(data, n);
    if constexpr (kSame) {
      search_space_ = build_space_;
    } else {
      search_space_ = std::make_shared<SearchSpace>(capacity_, dim_, metric_);
      search_space_->fit(data, n);
    }
    auto builder = std::make_shared<HNSWBuilder<BuildSpace>>(build_space_, max_nbrs_, efc);
    graph_ = builder->build_graph(threads);
    make_jobs();
  }

  // NSGBuilder::build_graph (include/index/graph/nsg/nsg_builder.hpp:97-161): NN-Descent + link + tree_grow
  void fit_nsg(float *data, uint32_t n, uint32_t efc) override {
    build_space_ = std::make_shared<BuildSpace>(capacity_, dim_, metric_);
    build_space_->fit(data, n);
    if constexpr (kSame) {
      search_space_ = build_space_;
    } else {
      search_space_ = std::make_shared<SearchSpace>(capacity_, dim_, metric_);
      search_space_->fit(data, n);
    }
    auto builder = std::make_shared<NSGBuilder<BuildSpace>>(build_space_, max_nbrs_, efc);
    graph_ = builder->build_graph();
    make_jobs();
  }
  // NndescentImpl::build_graph (include/index/graph/knng/nndescent.hpp:142-176): ids [n, k]
  void knng(float *data, uint32_t n, uint32_t k, uint32_t *out) override {
    auto sp = std::make_shared<BuildSpace>(capacity_, dim_, metric_);
    sp->fit(data, n);
    NndescentImpl<BuildSpace> nnd(sp, k);
    auto g = nnd.build_graph();
    for (uint32_t i = 0; i < n; i++)
      for (uint32_t j = 0; j < k; j++) out[size_t(i) * k + j] = g->at(i, j);
  }

  template <typename QC>
  static void rerank(std::vector<uint32_t> &src, uint32_t *dst, float *dist, QC qc, uint32_t ef,
                     uint32_t topk) {
    std::priority_queue<std::pair<float, uint32_t>, std::vector<std::pair<float, uint32_t>>,
                        std::greater<>>
        pq;
    for (size_t i = 0; i < ef; i++) pq.push({qc(src[i]), src[i]});
    for (size_t i = 0; i < topk; i++) {
      if (dist != nullptr) dist[i] = pq.top().first;
      dst[i] = pq.top().second;
      pq.pop();
    }
  }

  void search(float *q, uint32_t topk, uint32_t ef, uint32_t *ids) override {
    std::vector<uint32_t> res_pool(ef);
    search_job_->search_solo(q, topk, res_pool.data(), ef);
    if constexpr (kSame) {
      std::copy(res_pool.begin(), res_pool.begin() + topk, ids);
    } else {
      rerank(res_pool, ids, nullptr, build_space_->get_query_computer(q), ef, topk);
    }
  }

  // GraphSearchJob::search_solo_updated (graph_search_job.hpp:373-417): the search that consults the update
  // context (removed vertices expand through their saved neighbour lists); search-space order, no re-rank
  void search_updated(float *q, uint32_t topk, uint32_t ef, uint32_t *ids) override {
    search_job_->search_solo_updated(q, topk, ids, ef);
  }

  void finish(std::vector<std::vector<uint32_t>> &res_pool, float *q, uint32_t nq, uint32_t topk,
              uint32_t ef, uint32_t *ids) {
    if constexpr (kSame) {
      for (size_t i = 0; i < nq; i++)
        std::copy(res_pool[i].begin(), res_pool[i].begin() + topk, ids + i * topk);
    } else {
      for (size_t i = 0; i < nq; i++)
        rerank(res_pool[i], ids + i * topk, nullptr,
               build_space_->get_query_computer(q + i * dim_), ef, topk);
    }
  }

  void batch_search(float *q, uint32_t nq, uint32_t topk, uint32_t ef, uint32_t threads,
                    uint32_t *ids) override {
    std::vector<std::vector<uint32_t>> res_pool(nq, std::vector<uint32_t>(ef));
    std::vector<CpuID> worker_cpus;
    std::vector<coro::task<>> coros;
    coros.reserve(nq);
    for (uint32_t i = 0; i < threads; i++) worker_cpus.push_back(i);
    auto scheduler = std::make_shared<Scheduler>(worker_cpus);
    for (uint32_t i = 0; i < nq; i++) {
      coros.emplace_back(search_job_->search(q + size_t(i) * dim_, topk, res_pool[i].data(), ef));
      scheduler->schedule(coros.back().handle());
    }
    scheduler->begin();
    scheduler->join();
    finish(res_pool, q, nq, topk, ef, ids);
  }

  // tests/executor/search_test.cpp:147-174 pattern: std::thread x search_solo.
  void batch_search_solo(float *q, uint32_t nq, uint32_t topk, uint32_t ef, uint32_t threads,
                         uint32_t *ids) override {
    std::vector<std::vector<uint32_t>> res_pool(nq, std::vector<uint32_t>(ef));
    std::vector<std::thread> th;
    for (uint32_t t = 0; t < threads; t++) {
      th.emplace_back([&, t] {
        for (uint32_t i = t; i < nq; i += threads)
          search_job_->search_solo(q + size_t(i) * dim_, topk, res_pool[i].data(), ef);
      });
    }
    for (auto &x : th) x.join();
    finish(res_pool, q, nq, topk, ef, ids);
  }

  void batch_search_with_distance(float *q, uint32_t nq, uint32_t topk, uint32_t ef,
                                  uint32_t threads, uint32_t *ids, float *dist,
                                  int *has_dist) override {
    // index.hpp:372-418 (Linux branch): distances only when search space == build space.
    std::vector<std::vector<uint32_t>> res_pool(nq, std::vector<uint32_t>(ef));
    std::vector<std::vector<float>> dist_pool;
    std::vector<CpuID> worker_cpus;
    std::vector<coro::task<>> coros;
    coros.reserve(nq);
    for (uint32_t i = 0; i < threads; i++) worker_cpus.push_back(i);
    auto scheduler = std::make_shared<Scheduler>(worker_cpus);
    if constexpr (kSame) dist_pool.assign(nq, std::vector<float>(ef));
    for (uint32_t i = 0; i < nq; i++) {
      float *cur = q + size_t(i) * dim_;
      if constexpr (kSame) {
        coros.emplace_back(
            search_job_->search(cur, topk, res_pool[i].data(), dist_pool[i].data(), ef));
      } else {
        coros.emplace_back(search_job_->search(cur, topk, res_pool[i].data(), ef));
      }
      scheduler->schedule(coros.back().handle());
    }
    scheduler->begin();
    scheduler->join();
    for (size_t i = 0; i < nq; i++)
      std::copy(res_pool[i].begin(), res_pool[i].begin() + topk, ids + i * topk);
    *has_dist = kSame ? 1 : 0;
    if constexpr (kSame) {
      for (size_t i = 0; i < nq; i++)
        std::copy(dist_pool[i].begin(), dist_pool[i].begin() + topk, dist + i * topk);
    }
  }

  // Replays search_solo (graph_search_job.hpp:302-335) with counters; returns the whole pool.
  // stats = {level-0 distance calls, hops, overlay distance calls}
  void traverse(float *q, uint32_t ef, uint32_t *pool_ids, float *pool_dist, uint32_t *pool_size,
                uint64_t *stats) override {
    auto qc = search_space_->get_query_computer(q);
    LinearPool<float, uint32_t> pool(search_space_->get_data_num(), ef);
    uint64_t ndc = 0, hops = 0, ov = 0;
    if (graph_->overlay_graph_) {
      auto &o = *graph_->overlay_graph_;
      uint32_t u = o.ep_;
      float cd = qc(u);
      ov++;
      for (int l = o.levels_[u]; l > 0; --l) {
        bool ch = true;
        while (ch) {
          ch = false;
          auto *lst = o.edges(l, u);
          for (uint32_t j = 0; j < o.max_nbrs_ && lst[j] != static_cast<uint32_t>(-1); ++j) {
            float d = qc(lst[j]);
            ov++;
            if (d < cd) {
              cd = d;
              u = lst[j];
              ch = true;
            }
          }
        }
      }
      pool.insert(u, cd);
      pool.vis_.set(u);
    } else {
      graph_->initialize_search(pool, qc);
    }
    while (pool.has_next()) {
      auto u = pool.pop();
      hops++;
      for (uint32_t j = 0; j < graph_->max_nbrs_; ++j) {
        auto v = graph_->at(u, j);
        if (v == static_cast<uint32_t>(-1)) break;
        if (pool.vis_.get(v)) continue;
        pool.vis_.set(v);
        ndc++;
        pool.insert(v, qc(v));
      }
    }
    *pool_size = pool.size();
    for (size_t i = 0; i < pool.size(); i++) {
      pool_ids[i] = pool.id(i);
      pool_dist[i] = pool.dist(i);
    }
    stats[0] = ndc;
    stats[1] = hops;
    stats[2] = ov;
  }

  uint32_t insert(float *v, uint32_t ef) override { return update_job_->insert_and_update(v, ef); }
  void remove(uint32_t id) override { update_job_->remove(id); }

  void save(const char *ip, const char *dp, const char *qp) override {
    graph_->save(ip);
    if (dp != nullptr && dp[0] != 0) build_space_->save(dp);
    if (qp != nullptr && qp[0] != 0) search_space_->save(qp);
  }

  void load(const char *ip, const char *dp, const char *qp) override {
    graph_ = std::make_shared<Graph<float, uint32_t>>();
    graph_->load(ip);
    if (dp != nullptr && dp[0] != 0) {
      build_space_ = std::make_shared<BuildSpace>();
      build_space_->load(dp);
      build_space_->set_metric_function();
    }
    if constexpr (kSame) {
      search_space_ = build_space_;
    } else {
      search_space_ = std::make_shared<SearchSpace>();
      search_space_->load(qp);
      search_space_->set_metric_function();
    }
    dim_ = build_space_->dim_;
    make_jobs();
  }

  Graph<float, uint32_t> *graph() override { return graph_.get(); }
  RawSpace<> *raw() override { return build_space_.get(); }
  const uint8_t *code_row(uint32_t id) override {
    if constexpr (kSame) {
      return nullptr;
    } else {
      return search_space_->get_data_by_id(id);
    }
  }
  int quant_minmax(float *mn, float *mx) override {
    if constexpr (kSame) {
      return -1;
    } else {
      auto qz = search_space_->get_quantizer();
      std::memcpy(mn, qz.get_min(), dim_ * sizeof(float));
      std::memcpy(mx, qz.get_max(), dim_ * sizeof(float));
      return 0;
    }
  }
  float search_space_distance(uint32_t i, uint32_t j) override {
    return search_space_->get_distance(i, j);
  }
};

template <typename F>
int guarded(F &&f) {
  try {
    f();
    return 0;
  } catch (const std::exception &e) {
    g_err = e.what();
    return -1;
  }
}

}  // namespace

extern "C" {

const char *ref_last_error() { return g_err.c_str(); }
const char *ref_simd_level() { return simd::get_simd_level_name(); }
void ref_set_log_level(int lvl) { spdlog::set_level(static_cast<spdlog::level::level_enum>(lvl)); }

// metric: 0 L2, 1 IP, 2 COS (utils/metric_type.hpp:27-32). quant: 0 none, 1 sq8, 2 sq4.
void *ref_index_create(int metric, int quant, uint32_t capacity, uint32_t dim) {
  spdlog::set_level(spdlog::level::warn);
  auto m = static_cast<MetricType>(metric);
  RefIndexBase *p = nullptr;
  if (quant == 0) p = new RefIndex<RawSpace<>>(m, capacity);
  if (quant == 1) p = new RefIndex<SQ8Space<>>(m, capacity);
  if (quant == 2) p = new RefIndex<SQ4Space<>>(m, capacity);
  if (p != nullptr) p->dim_ = dim;
  return p;
}
void ref_index_destroy(void *h) { delete static_cast<RefIndexBase *>(h); }

int ref_index_fit(void *h, float *data, uint32_t n, uint32_t efc, uint32_t threads) {
  return guarded([&] { static_cast<RefIndexBase *>(h)->fit(data, n, efc, threads); });
}
int ref_index_fit_nsg(void *h, float *data, uint32_t n, uint32_t efc) {
  return guarded([&] { static_cast<RefIndexBase *>(h)->fit_nsg(data, n, efc); });
}
int ref_index_knng(void *h, float *data, uint32_t n, uint32_t k, uint32_t *out) {
  return guarded([&] { static_cast<RefIndexBase *>(h)->knng(data, n, k, out); });
}
int ref_index_search(void *h, float *q, uint32_t topk, uint32_t ef, uint32_t *ids) {
  return guarded([&] { static_cast<RefIndexBase *>(h)->search(q, topk, ef, ids); });
}
int ref_index_search_updated(void *h, float *q, uint32_t topk, uint32_t ef, uint32_t *ids) {
  return guarded([&] { static_cast<RefIndexBase *>(h)->search_updated(q, topk, ef, ids); });
}
int ref_index_batch_search(void *h, float *q, uint32_t nq, uint32_t topk, uint32_t ef,
                           uint32_t threads, uint32_t *ids) {
  return guarded([&] { static_cast<RefIndexBase *>(h)->batch_search(q, nq, topk, ef, threads, ids); });
}
int ref_index_batch_search_solo(void *h, float *q, uint32_t nq, uint32_t topk, uint32_t ef,
                                uint32_t threads, uint32_t *ids) {
  return guarded(
      [&] { static_cast<RefIndexBase *>(h)->batch_search_solo(q, nq, topk, ef, threads, ids); });
}
int ref_index_batch_search_with_distance(void *h, float *q, uint32_t nq, uint32_t topk,
                                         uint32_t ef, uint32_t threads, uint32_t *ids, float *dist,
                                         int *has_dist) {
  return guarded([&] {
    static_cast<RefIndexBase *>(h)->batch_search_with_distance(q, nq, topk, ef, threads, ids, dist,
                                                               has_dist);
  });
}
int ref_index_traverse(void *h, float *q, uint32_t ef, uint32_t *pool_ids, float *pool_dist,
                       uint32_t *pool_size, uint64_t *stats) {
  return guarded(
      [&] { static_cast<RefIndexBase *>(h)->traverse(q, ef, pool_ids, pool_dist, pool_size, stats); });
}
int64_t ref_index_insert(void *h, float *v, uint32_t ef) {
  int64_t out = -2;
  int rc = guarded([&] { out = static_cast<RefIndexBase *>(h)->insert(v, ef); });
  return rc == 0 ? out : -2;
}
int ref_index_remove(void *h, uint32_t id) {
  return guarded([&] { static_cast<RefIndexBase *>(h)->remove(id); });
}
int ref_index_save(void *h, const char *ip, const char *dp, const char *qp) {
  return guarded([&] { static_cast<RefIndexBase *>(h)->save(ip, dp, qp); });
}
int ref_index_load(void *h, const char *ip, const char *dp, const char *qp) {
  return guarded([&] { static_cast<RefIndexBase *>(h)->load(ip, dp, qp); });
}

// graph accessors -----------------------------------------------------------------------------
// info = {max_nodes, max_nbrs, ep, has_overlay, graph rows}
int ref_index_graph_info(void *h, uint32_t *info) {
  auto *b = static_cast<RefIndexBase *>(h);
  auto *g = b->graph();
  if (g == nullptr) return -1;
  info[0] = g->max_nodes_;
  info[1] = g->max_nbrs_;
  info[2] = g->get_ep();
  info[3] = g->overlay_graph_ ? 1 : 0;
  info[4] = static_cast<uint32_t>(g->data_storage_.pos_);  // graph rows (== search-space rows, also after inserts)
  return 0;
}
int ref_index_graph_rows(void *h, uint32_t first, uint32_t count, uint32_t *out) {
  auto *g = static_cast<RefIndexBase *>(h)->graph();
  for (uint32_t i = 0; i < count; i++)
    std::memcpy(out + size_t(i) * g->max_nbrs_, g->edges(first + i), g->max_nbrs_ * 4);
  return 0;
}
int ref_index_overlay_levels(void *h, uint32_t count, uint32_t *levels) {
  auto *g = static_cast<RefIndexBase *>(h)->graph();
  if (!g->overlay_graph_) return -1;
  for (uint32_t i = 0; i < count; i++) levels[i] = g->overlay_graph_->levels_[i];
  return 0;
}
// copies levels_[id] * max_nbrs ids
int ref_index_overlay_list(void *h, uint32_t id, uint32_t *out) {
  auto *g = static_cast<RefIndexBase *>(h)->graph();
  if (!g->overlay_graph_) return -1;
  auto &l = g->overlay_graph_->lists_[id];
  std::memcpy(out, l.data(), l.size() * 4);
  return static_cast<int>(l.size());
}
int ref_index_raw_row(void *h, uint32_t id, float *out) {
  auto *b = static_cast<RefIndexBase *>(h);
  std::memcpy(out, b->raw()->get_data_by_id(id), b->dim_ * sizeof(float));
  return 0;
}
int ref_index_code_row(void *h, uint32_t id, uint8_t *out, uint32_t nbytes) {
  auto *p = static_cast<RefIndexBase *>(h)->code_row(id);
  if (p == nullptr) return -1;
  std::memcpy(out, p, nbytes);
  return 0;
}
int ref_index_quant_minmax(void *h, float *mn, float *mx) {
  return static_cast<RefIndexBase *>(h)->quant_minmax(mn, mx);
}
float ref_index_space_distance(void *h, uint32_t i, uint32_t j) {
  return static_cast<RefIndexBase *>(h)->search_space_distance(i, j);
}

// kernels (include/simd/distance_l2.ipp, distance_ip.ipp; runtime-dispatched variants) ----------
float ref_l2_sqr(const float *x, const float *y, size_t d) { return simd::l2_sqr<float, float>(x, y, d); }
float ref_ip_sqr(const float *x, const float *y, size_t d) { return simd::ip_sqr<float, float>(x, y, d); }
float ref_l2_sqr_sq8(const uint8_t *x, const uint8_t *y, size_t d, const float *mn, const float *mx) {
  return simd::l2_sqr_sq8<float, float>(x, y, d, mn, mx);
}
float ref_ip_sqr_sq8(const uint8_t *x, const uint8_t *y, size_t d, const float *mn, const float *mx) {
  return simd::ip_sqr_sq8<float, float>(x, y, d, mn, mx);
}
float ref_l2_sqr_sq4(const uint8_t *x, const uint8_t *y, size_t d, const float *mn, const float *mx) {
  return simd::l2_sqr_sq4<float, float>(x, y, d, mn, mx);
}
float ref_ip_sqr_sq4(const uint8_t *x, const uint8_t *y, size_t d, const float *mn, const float *mx) {
  return simd::ip_sqr_sq4<float, float>(x, y, d, mn, mx);
}
// generic (scalar) variants, for pinning the restatement's summation order
float ref_l2_sqr_generic(const float *x, const float *y, size_t d) { return simd::l2_sqr_generic(x, y, d); }
float ref_ip_sqr_generic(const float *x, const float *y, size_t d) { return simd::ip_sqr_generic(x, y, d); }
float ref_l2_sqr_sq8_generic(const uint8_t *x, const uint8_t *y, size_t d, const float *mn, const float *mx) {
  return simd::l2_sqr_sq8_generic(x, y, d, mn, mx);
}
float ref_ip_sqr_sq8_generic(const uint8_t *x, const uint8_t *y, size_t d, const float *mn, const float *mx) {
  return simd::ip_sqr_sq8_generic(x, y, d, mn, mx);
}
float ref_l2_sqr_sq4_generic(const uint8_t *x, const uint8_t *y, size_t d, const float *mn, const float *mx) {
  return simd::l2_sqr_sq4_generic(x, y, d, mn, mx);
}
float ref_ip_sqr_sq4_generic(const uint8_t *x, const uint8_t *y, size_t d, const float *mn, const float *mx) {
  return simd::ip_sqr_sq4_generic(x, y, d, mn, mx);
}

// quantizers (include/space/quant/sq8.hpp:99-143, sq4.hpp:99-149) -------------------------------
int ref_sq8_fit_encode(const float *data, size_t n, uint32_t dim, float *mn, float *mx, uint8_t *codes) {
  SQ8Quantizer<float> qz(dim);
  qz.fit(data, n);
  std::memcpy(mn, qz.get_min(), dim * 4);
  std::memcpy(mx, qz.get_max(), dim * 4);
  for (size_t i = 0; i < n; i++) qz.encode(data + i * dim, codes + i * dim);
  return 0;
}
int ref_sq4_fit_encode(const float *data, size_t n, uint32_t dim, float *mn, float *mx, uint8_t *codes) {
  SQ4Quantizer<float> qz(dim);
  qz.fit(data, n);
  std::memcpy(mn, qz.get_min(), dim * 4);
  std::memcpy(mx, qz.get_max(), dim * 4);
  size_t rb = (dim + 1) / 2;
  for (size_t i = 0; i < n; i++) qz.encode(data + i * dim, codes + i * rb);
  return 0;
}
uint8_t ref_sq8_quantize(float v, float mn, float mx) { return SQ8Quantizer<float>(1).quantize(v, mn, mx); }
uint8_t ref_sq4_quantize(float v, float mn, float mx) { return SQ4Quantizer<float>(1).quantize(v, mn, mx); }
void ref_normalize(float *v, size_t d) { normalize(v, d); }

// brute force (include/utils/evaluate.hpp:30-62), one thread per slice of queries ----------------
int ref_find_exact_gt(const float *q, uint32_t nq, const float *data, uint32_t n, uint32_t dim,
                      uint32_t topk, uint32_t threads, uint32_t *ids) {
  std::vector<float> base(data, data + size_t(n) * dim);
  std::vector<std::thread> th;
  for (uint32_t t = 0; t < threads; t++) {
    th.emplace_back([&, t] {
      for (uint32_t i = t; i < nq; i += threads) {
        std::vector<float> one(q + size_t(i) * dim, q + size_t(i + 1) * dim);
        auto r = find_exact_gt<float, float, uint32_t>(one, base, dim, topk);
        std::copy(r.begin(), r.end(), ids + size_t(i) * topk);
      }
    });
  }
  for (auto &x : th) x.join();
  return 0;
}

// LinearPool replay for known-answer tests (tests/utils/query_utils_test.cpp) ---------------------
// ops: for each i, if op_ids[i] == 0xFFFFFFFF -> pop (result appended to popped), else insert.
int ref_linear_pool_replay(uint32_t n, int capacity, const uint32_t *op_ids, const float *op_dist,
                           uint32_t nops, uint32_t *popped, uint32_t *npopped, uint32_t *final_ids,
                           float *final_dist, uint32_t *final_size) {
  LinearPool<float, uint32_t> pool(n, capacity);
  uint32_t np = 0;
  for (uint32_t i = 0; i < nops; i++) {
    if (op_ids[i] == 0xFFFFFFFFu) {
      popped[np++] = pool.pop();
    } else {
      pool.insert(op_ids[i], op_dist[i]);
    }
  }
  *npopped = np;
  *final_size = pool.size();
  for (size_t i = 0; i < pool.size(); i++) {
    final_ids[i] = pool.id(i);
    final_dist[i] = pool.dist(i);
  }
  return 0;
}

}  // extern "C"
