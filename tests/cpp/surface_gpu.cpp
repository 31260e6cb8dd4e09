// End-to-end through the C++ surface on a GPU: fit a DeviceSpace, build on the device into a host graph
// structure, re-upload that graph into a second space, search both, print the ids.
#include <cmath>
#include <cstdio>
#include <random>

#include "alaya_b200.hpp"
#include "mock_graph.hpp"

int main() {
  const uint32_t n = 4000, d = 32, nq = 50, k = 10, ef = 64;
  std::mt19937 rng(7);
  std::normal_distribution<float> g;
  std::vector<float> base(size_t(n) * d), q(size_t(nq) * d);
  for (auto &x : base) x = g(rng);
  for (auto &x : q) x = g(rng);
  using Space = alaya_b200::DeviceSpace<>;
  auto space = std::make_shared<Space>(n, d, AB_METRIC_L2);
  space->fit(base.data(), n);
  alaya_b200::DeviceGraphBuilder<Space, MockGraph> builder(space, 32, 100);
  auto graph = builder.build_graph();
  alaya_b200::DeviceSearchJob<Space> job(space);
  std::vector<uint32_t> ids(size_t(nq) * k), ids2(size_t(nq) * k), one(k);
  std::vector<float> dist(k);
  job.batch_search(q.data(), nq, k, ids.data(), ef);
  // the same graph through the host structure into a fresh space must give the same answers
  auto space2 = std::make_shared<Space>(n, d, AB_METRIC_L2);
  space2->fit(base.data(), n);
  alaya_b200::DeviceSearchJob<Space> job2(space2);
  job2.set_graph(*graph);
  job2.batch_search(q.data(), nq, k, ids2.data(), ef);
  job.search_solo(q.data(), k, one.data(), dist.data(), ef);
  int same = ids == ids2, solo_same = 1, sorted = 1;
  for (uint32_t i = 0; i < k; i++) solo_same &= one[i] == ids[i];
  for (uint32_t i = 1; i < k; i++) sorted &= dist[i - 1] <= dist[i];
  const float d01 = space->get_distance(0, 1);
  float chk = 0;
  for (uint32_t j = 0; j < d; j++) chk += (base[j] - base[d + j]) * (base[j] - base[d + j]);
  std::printf("same=%d solo_same=%d sorted=%d dist_ok=%d n=%u\n", same, solo_same, sorted,
              std::abs(d01 - chk) <= 1e-4f * chk, unsigned(space->get_data_num()));
  for (uint32_t i = 0; i < nq * k; i++) std::printf("%u ", ids[i]);
  std::printf("\n");
  return 0;
}
