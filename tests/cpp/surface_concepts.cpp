// Checks the device-backed classes against the reference's own concepts (needs the reference checkout).
#include "index/graph/graph.hpp"
#include "index/graph/graph_concepts.hpp"
#include "space/space_concepts.hpp"

#include "alaya_b200.hpp"

using Space = alaya_b200::DeviceSpace<float, float, uint32_t>;
using Builder = alaya_b200::DeviceGraphBuilder<Space, alaya::Graph<float, uint32_t>>;
static_assert(alaya::Space<Space>, "DeviceSpace must satisfy alaya::Space (space_concepts.hpp:50-73)");
static_assert(alaya::GraphBuilder<Builder>, "DeviceGraphBuilder must satisfy alaya::GraphBuilder (graph_concepts.hpp:40-49)");
template class alaya_b200::DeviceGraphBuilder<Space, alaya::Graph<float, uint32_t>>;
using NsgBuilder = alaya_b200::DeviceNSGBuilder<Space, alaya::Graph<float, uint32_t>>;
static_assert(alaya::GraphBuilder<NsgBuilder>, "DeviceNSGBuilder must satisfy alaya::GraphBuilder");
template class alaya_b200::DeviceNSGBuilder<Space, alaya::Graph<float, uint32_t>>;
template void alaya_b200::DeviceSearchJob<Space>::set_graph<alaya::Graph<float, uint32_t>>(const alaya::Graph<float, uint32_t> &);
int main() { return 0; }
