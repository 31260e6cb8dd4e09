// Test infrastructure only (oracle/_ref build). The real include/space/rabitq_space.hpp needs
// Eigen3 + OpenMP (absent here) and RaBitQ is outside the hot path. This stub
// supplies just the names include/executor/jobs/graph_search_job.hpp refers to.
#pragma once
#include <cstddef>
#include <cstdint>
#include <type_traits>

namespace alaya {
template <typename DataType = float, typename DistanceType = float, typename IDType = uint32_t>
class RaBitQSpace {
 public:
  constexpr static size_t kDegreeBound = 32;
};
template <typename T>
struct is_rabitq_space : std::false_type {};
template <typename T>
inline constexpr bool is_rabitq_space_v = is_rabitq_space<T>::value;
}  // namespace alaya
