// Launcher of the traversal kernel over SQ codes, instantiated once per quantisation in its own translation unit.
#pragma once
#include "search_kernel.cuh"

namespace ab {

namespace {

template <int QUANT, bool IP>
void launch_search_sq_t(const SearchPlan &p, const SearchArgs &a, cudaStream_t s) {
  // three candidates in flight, at most 4 chunks per lane; longer code rows take the generic kernel
  if (a.nchunks <= 16) launch_one<QUANT, IP, 4, 4, false>(p, a, s);
  else if (a.nchunks <= 32) launch_one<QUANT, IP, 8, 4, false>(p, a, s);
  else if (a.nchunks <= 64) launch_one<QUANT, IP, 16, 4, false>(p, a, s);
  else if (a.nchunks <= 128) launch_one<QUANT, IP, 32, 4, false>(p, a, s);
  else launch_one<QUANT, IP, 32, 0, false>(p, a, s);
  AB_CUDA(cudaGetLastError());
  count_launch();
}

}  // namespace

}  // namespace ab
