// Group commit for concurrent Search callers (reference: one Search per request thread on one model,
// tests/test.h:1033-1062, search/gamma_engine.cc:74-97).  Host-only logic, no CUDA: capi.cu instantiates it with the
// IVFPQ request type, tests/coalesce_stress.cc with a fake one under ThreadSanitizer.
//
// Up to `slots` batches are in flight at once (each on its own search context), so the device always has the next batch
// queued behind the running ones.  A caller that arrives while that many are in flight waits, and the next batch to
// start takes the waiting requests with the same parameters along.  A batch that starts while the device is busy anyway
// first waits (at most wait_us, spinning) until its share of the recent callers has arrived: threads released by one
// batch come back within microseconds of one another and would otherwise each start a batch of their own.  A lone caller,
// or any caller that finds the device idle, starts at once.
#pragma once
#include <algorithm>
#include <chrono>
#include <condition_variable>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

namespace gb {

struct CoalescePolicy {
  int slots = 3;           // batches in flight (>= 1)
  int max_queries = 2048;  // queries per batch
  int wait_us = 60;        // gathering window of a batch that starts while another one runs (0 = none)
  int balance = 1;         // a batch takes at most (recent callers / slots) requests
};

// Req: `int n` (queries of the request) and `bool same(const Req &) const` (may travel in one batch).
template <typename Req>
class Coalescer {
 public:
  // Blocks until `me` has been executed — by this thread as the leader of a batch (run(group, err) is called once with
  // group[0] == &me and returns the batch's code, `err` the message for the callers taken along) or by another caller's
  // batch.  Returns that batch's code; *err_out receives its message when this request travelled in someone else's.
  template <typename Run>
  int submit(Req &me, const CoalescePolicy &T, Run &&run, std::string *err_out) {
    Pending self;
    self.req = &me;
    std::vector<Pending *> grp;
    {
      std::unique_lock<std::mutex> g(mu_);
      auto &w = waiting_;
      w.push_back(&self);
      cv_.wait(g, [&] { return self.done || (!self.taken && leaders_ < T.slots); });
      if (self.done) {
        if (err_out) *err_out = self.err;
        return self.rc;
      }
      leaders_++;
      // from here on this request belongs to its own batch: out of the waiting list BEFORE the lock is dropped below,
      // or a second batch forming meanwhile would take it along as well
      w.erase(std::find(w.begin(), w.end(), &self));
      auto share = [&] {
        peak_ = std::max(peak_, inflight_ + 1 + (int)w.size());
        return std::max(1, (peak_ + T.slots - 1) / T.slots);
      };
      if (T.wait_us > 0 && running_ > 0) {
        const auto t0 = std::chrono::steady_clock::now();
        for (;;) {
          const int want = std::min(share(), std::max(1, peak_ - inflight_));
          if (1 + (int)w.size() >= want || running_ == 0) break;
          if (std::chrono::steady_clock::now() - t0 > std::chrono::microseconds(T.wait_us)) break;
          g.unlock();
          std::this_thread::yield();
          g.lock();
        }
      }
      const int max_reqs = T.balance ? share() : (1 << 30);
      int total = me.n;
      grp.push_back(&self);
      size_t keep = 0;
      for (size_t i = 0; i < w.size(); i++) {
        Pending *r = w[i];
        if ((int)grp.size() < max_reqs && r->req->same(me) && total + r->req->n <= T.max_queries) {
          grp.push_back(r);
          r->taken = true;
          total += r->req->n;
        } else {
          w[keep++] = r;
        }
      }
      w.resize(keep);
      peak_ = std::max(inflight_ + (int)grp.size() + (int)w.size(), peak_ - 1);
      running_++;
      inflight_ += (int)grp.size();
    }
    std::vector<Req *> reqs;
    reqs.reserve(grp.size());
    for (Pending *p : grp) reqs.push_back(p->req);
    std::string err;
    const int rc = run(reqs, err);
    {
      std::lock_guard<std::mutex> g(mu_);
      leaders_--;
      running_--;
      inflight_ -= (int)grp.size();
      for (size_t i = 1; i < grp.size(); i++) {  // a follower may return (and its Pending die) as soon as done is set
        grp[i]->rc = rc;
        grp[i]->err = err;
        grp[i]->done = true;
      }
    }
    cv_.notify_all();
    return rc;
  }

 private:
  struct Pending {
    Req *req = nullptr;
    int rc = 0;
    bool taken = false, done = false;  // travelling in another caller's batch / that batch has finished
    std::string err;
  };
  std::mutex mu_;
  std::condition_variable cv_;
  std::vector<Pending *> waiting_;
  int leaders_ = 0;   // batches being formed or running (<= slots)
  int running_ = 0;   // batches handed to run()
  int inflight_ = 0;  // callers travelling in them
  int peak_ = 0;      // recent number of concurrent callers (in flight + waiting), decays by one per batch
};

}  // namespace gb
