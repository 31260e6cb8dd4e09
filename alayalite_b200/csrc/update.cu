// Incremental insert / remove (placeholder until GraphUpdateJob's semantics land in this file).
#include "index.cuh"

namespace ab {

uint32_t index_insert(ab_index *ix, const float *vec, uint32_t ef) {
  (void)ix;
  (void)vec;
  (void)ef;
  throw Error(AB_ERR_UNSUPPORTED, "insert is not implemented yet");
}

void index_remove(ab_index *ix, uint32_t id) {
  (void)ix;
  (void)id;
  throw Error(AB_ERR_UNSUPPORTED, "remove is not implemented yet");
}

}  // namespace ab
