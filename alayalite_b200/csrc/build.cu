// GPU HNSW construction (placeholder until the batched builder lands in this file).
#include "index.cuh"

namespace ab {

void index_build_graph(ab_index *ix, uint32_t ef_construction) {
  (void)ix;
  (void)ef_construction;
  throw Error(AB_ERR_UNSUPPORTED, "GPU graph build is not implemented yet");
}

}  // namespace ab
