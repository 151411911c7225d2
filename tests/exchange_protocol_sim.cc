// Model of the multi-GPU result exchange protocol of gamma_b200/csrc/comm.cu (window buffers indexed by epoch % NBUF,
// per-peer epoch flags with release / acquire, immediate or deferred wait), ranks played by host threads.  The payload
// words are PLAIN memory, so ThreadSanitizer reports any overwrite of a window buffer that is not ordered after its
// consumer by the flag protocol — the property the comm.cu header argues for NBUF = 4.  Built and run by
// tests/test_coalesce_cpu.py::test_exchange_protocol_model.
//   argv: ranks epochs nbuf mode(0 immediate, 1 deferred, 2 mixed per rank and step)
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <thread>
#include <vector>

int main(int argc, char **argv) {
  const int R = argc > 1 ? atoi(argv[1]) : 4;
  const int E = argc > 2 ? atoi(argv[2]) : 2000;
  const int NBUF = argc > 3 ? atoi(argv[3]) : 4;
  const int mode = argc > 4 ? atoi(argv[4]) : 2;
  // window[r][buf][slot of rank s] and flag[r][buf][from rank s], as in gb200_comm::slot_of / flags_of
  std::vector<unsigned> window((size_t)R * NBUF * R, 0u);
  std::vector<std::atomic<unsigned>> flag((size_t)R * NBUF * R);
  for (auto &f : flag) f.store(0u);
  auto W = [&](int r, int b, int s) -> unsigned & { return window[((size_t)r * NBUF + b) * R + s]; };
  auto F = [&](int r, int b, int s) -> std::atomic<unsigned> & { return flag[((size_t)r * NBUF + b) * R + s]; };
  std::atomic<long long> bad{0};
  auto rank_main = [&](int me) {
    std::mt19937 rng(99 + me);
    auto jitter = [&] {
      const int k = (int)(rng() % 64);
      if (k < 4) std::this_thread::sleep_for(std::chrono::microseconds(1 + rng() % 200));
      else if (k < 24) std::this_thread::yield();
    };
    unsigned pending = 0;  // epoch whose gathered window this rank still has to consume (deferred form)
    for (unsigned e = 1; e <= (unsigned)E; e++) {
      jitter();  // the search
      const bool deferred = mode == 1 || (mode == 2 && (rng() & 1));
      const int buf = (int)(e % NBUF);
      W(me, buf, me) = e;  // own result, written in place
      for (int p = 0; p < R; p++) {  // push: payload, then release flag (exchange kernel / re-rank kernel tail)
        if (p == me) continue;
        W(p, buf, me) = e;
        F(p, buf, me).store(e, std::memory_order_release);
      }
      const unsigned we = deferred ? e - 1 : e;  // wait
      if (we)
        for (int p = 0; p < R; p++) {
          if (p == me) continue;
          while ((int)(F(me, (int)(we % NBUF), p).load(std::memory_order_acquire) - we) < 0) std::this_thread::yield();
        }
      jitter();
      // consume, in stream order (after this exchange's wait, before the next search): everything waited for so far
      for (unsigned c = pending ? pending : we; c && c <= we; c++)
        for (int p = 0; p < R; p++)
          if (W(me, (int)(c % NBUF), p) != c) bad++;
      pending = we + 1;
    }
    // gb200_comm_flush: wait for the last epoch and consume what is left
    for (int p = 0; p < R; p++) {
      if (p == me) continue;
      while ((int)(F(me, E % NBUF, p).load(std::memory_order_acquire) - (unsigned)E) < 0) std::this_thread::yield();
    }
    for (unsigned c = pending; c <= (unsigned)E; c++)
      for (int p = 0; p < R; p++)
        if (W(me, (int)(c % NBUF), p) != c) bad++;
  };
  std::vector<std::thread> th;
  for (int r = 0; r < R; r++) th.emplace_back(rank_main, r);
  for (auto &t : th) t.join();
  printf("{\"ranks\":%d,\"epochs\":%d,\"nbuf\":%d,\"mode\":%d,\"bad\":%lld}\n", R, E, NBUF, mode, bad.load());
  return bad.load() ? 1 : 0;
}
