// TEST INFRASTRUCTURE (oracle): stand-in for <spikes/thread_pool.hpp>
// (call sites: reference src/core/linear_form.hpp:11,28-53: thread_pool(n), size(),
// enqueue(f) -> std::future<void>). Runs each task on its own std::thread.
#ifndef ORACLE_SHIM_SPIKES_THREAD_POOL_HPP
#define ORACLE_SHIM_SPIKES_THREAD_POOL_HPP

#include <thread>
#include <future>
#include <functional>
#include <cstddef>

class thread_pool {
public:
  explicit thread_pool(std::size_t n) : n_(n ? n : 1) {}
  std::size_t size() const { return n_; }

  template<typename F>
  std::future<void> enqueue(F f) {
    return std::async(std::launch::async, f);
  }
private:
  std::size_t n_;
};

#endif
