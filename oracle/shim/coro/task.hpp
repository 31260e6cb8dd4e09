// Test infrastructure only (oracle/_ref build): stand-in for libcoro 0.14.1's coro::task<>,
// which is pinned in /root/reference/conanfile.py but not vendored. The reference hot path only
// needs a lazily-started, externally-resumed void task exposing .handle()
// (include/executor/jobs/graph_search_job.hpp:221, python/include/index.hpp:304-322).
#pragma once
#include <coroutine>
#include <exception>
#include <utility>

namespace coro {

template <typename T = void>
class task;

template <>
class task<void> {
 public:
  struct promise_type {
    auto get_return_object() -> task {
      return task{std::coroutine_handle<promise_type>::from_promise(*this)};
    }
    auto initial_suspend() noexcept -> std::suspend_always { return {}; }
    auto final_suspend() noexcept -> std::suspend_always { return {}; }
    void return_void() {}
    void unhandled_exception() { std::terminate(); }
  };

  explicit task(std::coroutine_handle<promise_type> h) : h_(h) {}
  task(task &&o) noexcept : h_(std::exchange(o.h_, nullptr)) {}
  task(const task &) = delete;
  auto operator=(const task &) -> task & = delete;
  ~task() {
    if (h_) h_.destroy();
  }
  auto handle() -> std::coroutine_handle<promise_type> { return h_; }

 private:
  std::coroutine_handle<promise_type> h_;
};

}  // namespace coro
