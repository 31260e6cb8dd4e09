// Stand-in with the member set of the reference's alaya::Graph / alaya::OverlayGraph
// (include/index/graph/graph.hpp:47-158, overlay_graph.hpp:36-90) that alaya_b200.hpp touches, so the
// header can be exercised where the reference checkout is absent (the GPU box).
#pragma once
#include <cstdint>
#include <memory>
#include <vector>

struct MockOverlay {
  uint32_t node_num_, max_nbrs_, ep_ = 0;
  std::vector<uint32_t> levels_;
  std::vector<std::vector<uint32_t>> lists_;
  MockOverlay(uint32_t n, uint32_t r) : node_num_(n), max_nbrs_(r), levels_(n), lists_(n) {}
  void set_level(uint32_t i, uint32_t l) { levels_[i] = l; }
};

struct MockGraph {
  using OverlayGraphType = MockOverlay;
  uint32_t max_nodes_, max_nbrs_;
  std::vector<uint32_t> rows_;
  uint32_t pos_ = 0;
  std::unique_ptr<MockOverlay> overlay_graph_;
  std::vector<uint32_t> eps_;
  MockGraph(uint32_t n, uint32_t r) : max_nodes_(n), max_nbrs_(r), rows_(size_t(n) * r, 0xFFFFFFFFu) {}
  uint32_t &at(uint32_t i, uint32_t j) { return rows_[size_t(i) * max_nbrs_ + j]; }
  uint32_t at(uint32_t i, uint32_t j) const { return rows_[size_t(i) * max_nbrs_ + j]; }
  uint32_t insert(const uint32_t *e) {
    for (uint32_t j = 0; j < max_nbrs_; j++) rows_[size_t(pos_) * max_nbrs_ + j] = e[j];
    return pos_++;
  }
  uint32_t get_ep() const { return overlay_graph_ ? overlay_graph_->ep_ : eps_[0]; }
};
