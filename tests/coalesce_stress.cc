// Stress test of the host-side batching policy of concurrent searches (gamma_b200/csrc/coalesce.h) with a fake device:
// built with -fsanitize=thread by tests/test_coalesce_cpu.py.  Checks, for several policies and caller mixes, that every
// request is executed exactly once, by a batch of compatible requests within the size limits, that results and errors
// reach the right caller, and that never more than `slots` batches run at once.
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <string>
#include <thread>
#include <vector>

#include "../gamma_b200/csrc/coalesce.h"

struct FakeReq {
  int n = 0;
  int key = 0;       // requests with equal keys may share a batch
  long long in = 0;  // "query"
  long long out = 0; // "result", written by whoever runs the batch
  int executed = 0;
  bool same(const FakeReq &o) const { return key == o.key; }
};

static void busy_us(int us) {
  const auto t0 = std::chrono::steady_clock::now();
  while (std::chrono::steady_clock::now() - t0 < std::chrono::microseconds(us)) {
  }
}

int main(int argc, char **argv) {
  const int threads = argc > 1 ? atoi(argv[1]) : 12;
  const int iters = argc > 2 ? atoi(argv[2]) : 400;
  long long failures = 0;
  struct Case {
    gb::CoalescePolicy pol;
    int keys;
  };
  std::vector<Case> cases;
  for (int slots : {1, 2, 3, 4})
    for (int wait_us : {0, 40})
      for (int balance : {0, 1}) {
        Case c;
        c.pol.slots = slots, c.pol.wait_us = wait_us, c.pol.balance = balance, c.pol.max_queries = 64;
        c.keys = 1 + (slots & 1) * 2;  // one parameter set, or three
        cases.push_back(c);
      }
  printf("[");
  for (size_t ci = 0; ci < cases.size(); ci++) {
    const Case &c = cases[ci];
    gb::Coalescer<FakeReq> co;
    std::atomic<int> running{0}, max_running{0};
    std::atomic<long long> batches{0}, batched_reqs{0}, bad{0};
    auto worker = [&](int t) {
      std::mt19937 rng(1234 + t);
      for (int it = 0; it < iters; it++) {
        FakeReq r;
        r.n = 1 + (int)(rng() % 12);
        r.key = (int)(rng() % c.keys);
        r.in = (long long)t * 1000003 + it;
        std::string err;
        const int rc = co.submit(
            r, c.pol,
            [&](const std::vector<FakeReq *> &grp, std::string &e) {
              const int now = ++running;
              int m = max_running.load();
              while (now > m && !max_running.compare_exchange_weak(m, now)) {
              }
              int total = 0;
              for (FakeReq *q : grp) {
                total += q->n;
                if (!q->same(*grp[0])) bad++;
                q->executed++;
                q->out = q->in * 3 + 1;
              }
              if (total > c.pol.max_queries && grp.size() > 1) bad++;
              if (grp[0] != &r) bad++;
              batches++;
              batched_reqs += (long long)grp.size();
              busy_us(30 + (int)(grp.size() * 5));
              --running;
              if (grp[0]->key == 2 && (grp[0]->in & 7) == 0) {  // a failing batch: every member must see code and message
                e = "boom " + std::to_string(grp[0]->in);
                return -3;
              }
              return 0;
            },
            &err);
        if (r.executed != 1 || r.out != r.in * 3 + 1) bad++;
        if (rc != 0 && rc != -3) bad++;
        if (rc == -3 && r.key != 2) bad++;
        if (rc == -3 && err.empty() == false && err.rfind("boom ", 0) != 0) bad++;
        if ((rng() & 3) == 0) busy_us((int)(rng() % 20));
      }
    };
    std::vector<std::thread> th;
    for (int t = 0; t < threads; t++) th.emplace_back(worker, t);
    for (auto &x : th) x.join();
    if (max_running.load() > c.pol.slots) bad++;
    if (batched_reqs.load() != (long long)threads * iters) bad++;
    failures += bad.load();
    printf("%s{\"slots\":%d,\"wait_us\":%d,\"balance\":%d,\"keys\":%d,\"batches\":%lld,\"requests\":%lld,\"max_running\":%d,\"bad\":%lld}",
           ci ? "," : "", c.pol.slots, c.pol.wait_us, c.pol.balance, c.keys, batches.load(), batched_reqs.load(),
           max_running.load(), bad.load());
  }
  printf("]\n");
  return failures ? 1 : 0;
}
