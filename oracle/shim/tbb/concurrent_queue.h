// Test-infrastructure shim (NOT product code): a minimal stand-in for
// tbb::concurrent_bounded_queue so the reference's index/retrieval_model.h
// (which holds one such queue member, retrieval_model.h:306) compiles in a
// container without TBB.  Only the members the reference touches exist.
#pragma once
#include <condition_variable>
#include <deque>
#include <mutex>

namespace tbb {
template <typename T>
class concurrent_bounded_queue {
 public:
  void push(const T &v) {
    std::lock_guard<std::mutex> g(mu_);
    q_.push_back(v);
    cv_.notify_one();
  }
  bool try_push(const T &v) {
    push(v);
    return true;
  }
  bool try_pop(T &out) {
    std::lock_guard<std::mutex> g(mu_);
    if (q_.empty()) return false;
    out = q_.front();
    q_.pop_front();
    return true;
  }
  void pop(T &out) {
    std::unique_lock<std::mutex> g(mu_);
    cv_.wait(g, [&] { return !q_.empty(); });
    out = q_.front();
    q_.pop_front();
  }
  long size() const {
    std::lock_guard<std::mutex> g(mu_);
    return (long)q_.size();
  }
  bool empty() const { return size() == 0; }
  void set_capacity(long) {}
  void clear() {
    std::lock_guard<std::mutex> g(mu_);
    q_.clear();
  }

 private:
  mutable std::mutex mu_;
  std::condition_variable cv_;
  std::deque<T> q_;
};
template <typename T>
using concurrent_queue = concurrent_bounded_queue<T>;
}  // namespace tbb
