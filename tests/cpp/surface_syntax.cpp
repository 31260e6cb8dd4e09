// Instantiates every template of include/alaya_b200.hpp (compile-only on CPU boxes).
#include "alaya_b200.hpp"
#include "mock_graph.hpp"

using Space = alaya_b200::DeviceSpace<>;
template class alaya_b200::DeviceSpace<float, float, uint32_t>;
template class alaya_b200::DeviceGraphBuilder<Space, MockGraph>;
template class alaya_b200::DeviceNSGBuilder<Space, MockGraph>;
template class alaya_b200::DeviceSearchJob<Space>;
template void alaya_b200::DeviceSearchJob<Space>::set_graph<MockGraph>(const MockGraph &);

int main() {
  // without a CUDA device the library must refuse loudly (no CPU fallback)
  try {
    Space s(100, 8, AB_METRIC_L2);
  } catch (const std::runtime_error &e) {
    return 42;
  }
  return 0;
}
