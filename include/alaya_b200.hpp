// alaya_b200.hpp -- header-only C++ layer over the C ABI (alaya_b200.h), shaped like the C++ plugin
// surface of the reference so that C++ callers written against it keep compiling:
//
//   alaya::Space concept            include/space/space_concepts.hpp:50-73      -> alaya_b200::DeviceSpace
//   alaya::GraphBuilder concept     include/index/graph/graph_concepts.hpp:40-49 -> alaya_b200::DeviceGraphBuilder
//   GraphSearchJob<Space> methods   include/executor/jobs/graph_search_job.hpp:221,260,302,337
//                                                                                -> alaya_b200::DeviceSearchJob
//   GraphUpdateJob insert / remove  include/executor/jobs/graph_update_job.hpp:66-103
//
// (paths relative to the reference checkout). Nothing here computes on the host: every method forwards
// to libalaya_b200.so and throws std::runtime_error(ab_last_error()) on failure, the way the reference
// throws. tests/test_cpp_surface.py checks the two concepts against the reference's own headers.
//
// Requires C++17 (C++20 to test the concepts). Link with -lalaya_b200.
#ifndef ALAYA_B200_HPP_
#define ALAYA_B200_HPP_

#include <cstddef>
#include <cstdint>
#include <memory>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <vector>

#include "alaya_b200.h"

namespace alaya_b200 {

inline void check(int rc) {
  if (rc != AB_OK) throw std::runtime_error(ab_last_error());
}

// Owning handle; the methods are PyIndexInterface's (python/include/index.hpp:508-587).
class DeviceIndex {
 public:
  DeviceIndex(int metric, int quant, uint32_t capacity, int index_type = AB_INDEX_HNSW, int device = -1) {
    ab_params p{index_type, quant, metric, capacity, 32u, device};
    ab_index *raw = nullptr;
    check(ab_index_create(&p, &raw));
    h_.reset(raw, ab_index_destroy);
  }
  ab_index *get() const { return h_.get(); }

  void fit(const float *data, uint64_t n, uint32_t dim, uint32_t ef_construction = 100) {
    check(ab_index_fit(get(), data, n, dim, ef_construction, 1));
  }
  void fit_spaces(const float *data, uint64_t n, uint32_t dim) { check(ab_index_fit_spaces(get(), data, n, dim)); }
  void build_graph(uint32_t ef_construction) { check(ab_index_build_graph(get(), ef_construction)); }
  void search(const float *query, uint32_t topk, uint32_t ef, uint32_t *ids) {
    check(ab_index_search(get(), query, topk, ef, ids));
  }
  void batch_search(const float *queries, uint64_t nq, uint32_t topk, uint32_t ef, uint32_t *ids) {
    check(ab_index_batch_search(get(), queries, nq, topk, ef, 1, ids));
  }
  // returns false when the space is quantised (the reference hands back an empty distance array there)
  bool batch_search_with_distance(const float *queries, uint64_t nq, uint32_t topk, uint32_t ef, uint32_t *ids,
                                  float *dist) {
    int32_t has = 0;
    check(ab_index_batch_search_with_distance(get(), queries, nq, topk, ef, 1, ids, dist, &has));
    return has != 0;
  }
  uint32_t insert(const float *vec, uint32_t ef) {
    uint32_t id = 0;
    check(ab_index_insert(get(), vec, ef, &id));
    return id;
  }
  void remove(uint32_t id) { check(ab_index_remove(get(), id)); }
  std::vector<float> get_data_by_id(uint32_t id) const {
    std::vector<float> out(ab_index_get_dim(get()));
    check(ab_index_get_data_by_id(get(), id, out.data()));
    return out;
  }
  uint32_t dim() const { return ab_index_get_dim(get()); }
  uint64_t size() const { return ab_index_get_size(get()); }
  void save(const std::string &index_path, const std::string &data_path, const std::string &quant_path = "") {
    check(ab_index_save(get(), index_path.c_str(), data_path.c_str(), quant_path.c_str()));
  }
  void load(const std::string &index_path, const std::string &data_path, const std::string &quant_path = "") {
    check(ab_index_load(get(), index_path.c_str(), data_path.c_str(), quant_path.c_str()));
  }

 private:
  std::shared_ptr<ab_index> h_;
};

// Function-pointer types of space_concepts.hpp:31-39
template <typename DataType, typename DistanceType>
using DistFunc = DistanceType (*)(const DataType *, const DataType *, size_t);

namespace detail {
template <int METRIC>
inline float pair_distance(const float *x, const float *y, size_t dim) {
  float out = 0.f;
  check(ab_distance_pairs(METRIC, AB_QUANT_NONE, x, y, 1, static_cast<uint32_t>(dim), nullptr, nullptr, &out));
  return out;
}
}  // namespace detail

// Device-resident vector space. Satisfies alaya::Space: the same member set as RawSpace
// (include/space/raw_space.hpp:47-330) as far as the concept and GraphSearchJob use it. Rows live in HBM;
// quant selects the *search* space of the index (raw / SQ8 / SQ4), the raw rows are always kept for
// the re-rank like PyIndex does (python/include/index.hpp:209-226).
template <typename DataType = float, typename DistanceType = float, typename IDType = uint32_t>
class DeviceSpace {
  static_assert(std::is_same<DataType, float>::value && std::is_same<DistanceType, float>::value,
                "the device path is fp32; convert other element types before fit() like dispatch.hpp does");

 public:
  using DataTypeAlias = DataType;
  using DistanceTypeAlias = DistanceType;
  using IDTypeAlias = IDType;
  using DistDataType = DataType;

  DeviceSpace(IDType capacity, size_t dim, int metric, int quant = AB_QUANT_NONE, int device = -1)
      : capacity_(capacity), dim_(dim), metric_(metric), index_(metric, quant, static_cast<uint32_t>(capacity),
                                                                AB_INDEX_HNSW, device) {}

  void fit(const DataType *data, IDType n) {  // RawSpace::fit, raw_space.hpp:118-153
    index_.fit_spaces(data, static_cast<uint64_t>(n), static_cast<uint32_t>(dim_));
    n_ = n;
  }
  void set_metric_function() {}  // the metric is fixed at construction; kept for the concept
  auto get_dist_func() const -> DistFunc<DataType, DistanceType> {
    return metric_ == AB_METRIC_L2 ? &detail::pair_distance<AB_METRIC_L2> : &detail::pair_distance<AB_METRIC_IP>;
  }
  auto get_distance(IDType i, IDType j) const -> DistanceType {  // raw_space.hpp:170-175
    const auto x = index_.get_data_by_id(static_cast<uint32_t>(i));
    const auto y = index_.get_data_by_id(static_cast<uint32_t>(j));
    return get_dist_func()(x.data(), y.data(), dim_);
  }
  auto get_dim() const -> size_t { return dim_; }
  auto get_data_size() const -> size_t { return dim_ * sizeof(DataType); }
  auto get_capacity() const -> IDType { return capacity_; }
  auto get_data_num() const -> IDType { return static_cast<IDType>(index_.size()); }
  auto get_data_by_id(IDType id) const -> std::vector<DataType> { return index_.get_data_by_id(static_cast<uint32_t>(id)); }
  auto insert(const DataType *vec, uint32_t ef = 100) -> IDType { return static_cast<IDType>(index_.insert(vec, ef)); }
  auto remove(IDType id) -> void { index_.remove(static_cast<uint32_t>(id)); }

  DeviceIndex &index() { return index_; }
  const DeviceIndex &index() const { return index_; }

 private:
  IDType capacity_;
  size_t dim_;
  int metric_;
  IDType n_ = 0;
  DeviceIndex index_;
};

// HNSW build on the device (replaces HNSWBuilder, include/index/graph/hnsw/hnsw_builder.hpp:54-194).
// GraphT is the reference's alaya::Graph<DataType, IDType> (or anything with the same members): the
// result is copied back into it so that host code keeps working on the structure it knows. The graph
// also stays resident on the device inside the space's index, ready for DeviceSearchJob.
template <class SpaceT, class GraphT>
class DeviceGraphBuilder {
 public:
  using DistanceSpaceTypeAlias = SpaceT;
  explicit DeviceGraphBuilder(std::shared_ptr<SpaceT> space, uint32_t max_nbrs = 32, uint32_t ef_construction = 200)
      : space_(std::move(space)), ef_construction_(ef_construction) {
    if (max_nbrs != 32) throw std::runtime_error("DeviceGraphBuilder: max_nbrs must be 32");
  }

  auto build_graph(uint32_t /*thread_num*/ = 1) -> std::unique_ptr<GraphT> {
    auto &ix = space_->index();
    ix.build_graph(ef_construction_);
    uint64_t info[5];
    check(ab_index_graph_info(ix.get(), info));
    const uint64_t n = info[0], R = info[1], lists_len = info[3];
    std::vector<uint32_t> adj(n * R), levels(n), lists(lists_len);
    check(ab_index_get_graph(ix.get(), adj.data(), levels.data(), lists.data()));
    // the members HNSWBuilder::build_graph fills (hnsw_builder.hpp:149-191)
    using Node = std::remove_reference_t<decltype(std::declval<GraphT &>().at(0, 0))>;
    auto g = std::make_unique<GraphT>(static_cast<Node>(space_->get_capacity()), static_cast<uint32_t>(R));
    std::vector<Node> row(R);
    for (uint64_t i = 0; i < n; i++) {
      for (uint64_t j = 0; j < R; j++) row[j] = static_cast<Node>(adj[i * R + j]);
      g->insert(row.data());
    }
    using Overlay = typename GraphT::OverlayGraphType;
    g->overlay_graph_ = std::make_unique<Overlay>(static_cast<Node>(n), static_cast<uint32_t>(R));
    g->overlay_graph_->ep_ = static_cast<Node>(info[2]);
    size_t off = 0;
    for (uint64_t i = 0; i < n; i++) {
      g->overlay_graph_->set_level(static_cast<Node>(i), levels[i]);
      if (levels[i] > 0) {
        auto &dst = g->overlay_graph_->lists_[i];
        dst.assign(lists.begin() + off, lists.begin() + off + size_t(levels[i]) * R);
        off += size_t(levels[i]) * R;
      }
    }
    return g;
  }

 private:
  std::shared_ptr<SpaceT> space_;
  uint32_t ef_construction_;
};

// NSGBuilder's shape (include/index/graph/nsg/nsg_builder.hpp:60-161): NN-Descent k-NN graph + MRNG link + reverse
// links + connectivity pass on the device (ab_index_build_graph_nsg). The result is the reference's own Graph with
// eps_ = {navigating node} and no overlay, and it stays resident in the space's index for DeviceSearchJob.
template <class SpaceT, class GraphT>
class DeviceNSGBuilder {
 public:
  using DistanceSpaceTypeAlias = SpaceT;
  explicit DeviceNSGBuilder(std::shared_ptr<SpaceT> space, uint32_t max_nbrs = 32, uint32_t ef_construction = 200)
      : space_(std::move(space)), ef_construction_(ef_construction) {
    if (max_nbrs != 32) throw std::runtime_error("DeviceNSGBuilder: max_nbrs must be 32");
  }
  uint32_t nndescent_iters_ = 10;  // NSGBuilder::nndescent_iters_

  auto build_graph(uint32_t /*thread_num*/ = 1) -> std::unique_ptr<GraphT> {
    auto &ix = space_->index();
    check(ab_index_build_graph_nsg(ix.get(), ef_construction_, nndescent_iters_));
    uint64_t info[5];
    check(ab_index_graph_info(ix.get(), info));
    const uint64_t n = info[0], R = info[1];
    std::vector<uint32_t> adj(n * R), levels(n), lists(1);
    check(ab_index_get_graph(ix.get(), adj.data(), levels.data(), lists.data()));
    using Node = std::remove_reference_t<decltype(std::declval<GraphT &>().at(0, 0))>;
    auto g = std::make_unique<GraphT>(static_cast<Node>(space_->get_capacity()), static_cast<uint32_t>(R));
    std::vector<Node> row(R);
    for (uint64_t i = 0; i < n; i++) {
      for (uint64_t j = 0; j < R; j++) row[j] = static_cast<Node>(adj[i * R + j]);
      g->insert(row.data());
    }
    g->eps_.push_back(static_cast<Node>(info[2]));  // nsg_builder.hpp:113
    return g;
  }

 private:
  std::shared_ptr<SpaceT> space_;
  uint32_t ef_construction_;
};

// Search over the graph resident in the space's index (replaces GraphSearchJob,
// include/executor/jobs/graph_search_job.hpp). The coroutine / Scheduler machinery has no counterpart:
// batch_search is one kernel launch over all queries.
template <class SpaceT>
class DeviceSearchJob {
 public:
  using DataType = typename SpaceT::DataTypeAlias;
  using DistanceType = typename SpaceT::DistanceTypeAlias;
  using IDType = typename SpaceT::IDTypeAlias;

  explicit DeviceSearchJob(std::shared_ptr<SpaceT> space) : space_(std::move(space)) {}

  // Upload a graph built elsewhere (e.g. by the reference's HNSWBuilder on the host).
  template <class GraphT>
  void set_graph(const GraphT &g) {
    const uint64_t n = space_->get_data_num();
    const uint32_t R = g.max_nbrs_;
    std::vector<uint32_t> adj(n * R), levels(n, 0), lists;
    for (uint64_t i = 0; i < n; i++)
      for (uint32_t j = 0; j < R; j++) adj[i * R + j] = static_cast<uint32_t>(g.at(static_cast<IDType>(i), j));
    uint32_t ep = g.get_ep();
    if (g.overlay_graph_) {
      for (uint64_t i = 0; i < n; i++) {
        levels[i] = g.overlay_graph_->levels_[i];
        const auto &l = g.overlay_graph_->lists_[i];
        lists.insert(lists.end(), l.begin(), l.begin() + size_t(levels[i]) * R);
      }
    }
    check(ab_index_set_graph(space_->index().get(), adj.data(), levels.data(), lists.data(), lists.size(), ep));
  }

  // graph_search_job.hpp:302-335
  void search_solo(const DataType *query, uint32_t k, IDType *ids, uint32_t ef) { batch_search(query, 1, k, ids, ef); }
  // graph_search_job.hpp:337-371
  void search_solo(const DataType *query, uint32_t k, IDType *ids, DistanceType *distances, uint32_t ef) {
    std::vector<uint32_t> tmp(k);
    space_->index().batch_search_with_distance(query, 1, k, ef, tmp.data(), distances);
    for (uint32_t i = 0; i < k; i++) ids[i] = static_cast<IDType>(tmp[i]);
  }
  // graph_search_job.hpp:373-417. The reference expands a removed vertex through the neighbour list saved at
  // removal time; on the device a removed vertex's adjacency row is frozen at that very list (updates skip invalid
  // rows like SequentialStorage::update does), so the ordinary traversal IS the updated one. Pool order, no re-rank.
  void search_solo_updated(const DataType *query, uint32_t k, IDType *ids, uint32_t ef) {
    std::vector<uint32_t> tmp(k);
    std::vector<float> dist(k);
    space_->index().batch_search_with_distance(query, 1, k, ef, tmp.data(), dist.data());
    for (uint32_t i = 0; i < k; i++) ids[i] = static_cast<IDType>(tmp[i]);
  }
  // PyIndex::batch_search (python/include/index.hpp:289-346) without the Python types
  void batch_search(const DataType *queries, uint64_t nq, uint32_t k, IDType *ids, uint32_t ef) {
    if constexpr (std::is_same<IDType, uint32_t>::value) {
      space_->index().batch_search(queries, nq, k, ef, ids);
    } else {
      std::vector<uint32_t> tmp(nq * k);
      space_->index().batch_search(queries, nq, k, ef, tmp.data());
      for (size_t i = 0; i < tmp.size(); i++) ids[i] = static_cast<IDType>(tmp[i]);
    }
  }

 private:
  std::shared_ptr<SpaceT> space_;
};

}  // namespace alaya_b200

#endif  // ALAYA_B200_HPP_
