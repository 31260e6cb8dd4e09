// beam_search_kernel over SQ8 codes (see search_kernel.cuh / search_sq.cuh)
#include "search_sq.cuh"

namespace ab {

void launch_search_sq8(bool ip, const SearchPlan &plan, const SearchArgs &args, cudaStream_t stream) {
  ip ? launch_search_sq_t<AB_QUANT_SQ8, true>(plan, args, stream) : launch_search_sq_t<AB_QUANT_SQ8, false>(plan, args, stream);
}

}  // namespace ab
