// Host model of the persistent mode's ticket ring queues (ygl_kernels.cu: ring_reserve / ring_publish / ring_take):
// the same protocol on std::atomic, driven by threads that play traversal, shading and light-pdf warps. Checks what the
// device code relies on: every lane is consumed exactly once per hop, tickets may run ahead of the producers, no
// entry is lost or duplicated, and the job terminates. Build: g++ -O2 -std=c++17 -pthread ring_stress.cpp
// usage: ring_stress [lanes] [hops_per_lane] [threads_per_role] [capacity_log2]
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <thread>
#include <vector>

struct Ring {
  std::vector<std::atomic<int>> slots;
  std::atomic<unsigned>         head{0}, tail{0};
  unsigned                      mask;
  explicit Ring(unsigned cap) : slots(cap), mask(cap - 1) {
    for (auto& s : slots) s.store(-1, std::memory_order_relaxed);
  }
};

static std::atomic<int> g_abort{0};

// producer: reserve an index, wait for the wrap-around guard, publish with release
static void ring_push(Ring& r, int value) {
  unsigned idx  = r.tail.fetch_add(1, std::memory_order_relaxed);
  auto&    slot = r.slots[idx & r.mask];
  long     spins = 0;
  while (slot.load(std::memory_order_relaxed) != -1)
    if (++spins > (1L << 28)) {
      g_abort.store(1);
      return;
    }
  slot.store(value, std::memory_order_release);
}
// consumer: take a ticket once, then poll the own slot without blocking
struct Ticket {
  unsigned idx  = 0;
  bool     held = false;
};
static bool ring_take(Ring& r, Ticket& t, int& entry) {
  if (!t.held) t.idx = r.head.fetch_add(1, std::memory_order_relaxed), t.held = true;
  auto& slot = r.slots[t.idx & r.mask];
  int   v    = slot.load(std::memory_order_acquire);
  if (v == -1) return false;
  slot.store(-1, std::memory_order_relaxed);
  t.held = false;
  entry  = v;
  return true;
}

int main(int argc, char** argv) {
  const int lanes   = argc > 1 ? atoi(argv[1]) : 20000;
  const int hops    = argc > 2 ? atoi(argv[2]) : 40;
  const int threads = argc > 3 ? atoi(argv[3]) : 4;
  const int caplog  = argc > 4 ? atoi(argv[4]) : 0;
  // capacity rule of run_persistent: lanes + every consumer that can hold a ticket, rounded up to a power of two
  unsigned cap = 64;
  while (cap < (unsigned)(lanes + 3 * threads * 32)) cap <<= 1;
  if (caplog) cap = 1u << caplog;
  Ring ext(cap), shade(cap), lpdf(cap);
  std::vector<std::atomic<int>> visits(lanes);   // hops done per lane
  std::vector<std::atomic<int>> in_flight(lanes);  // must be 0 or 1: a lane sits in one place at a time
  for (auto& v : visits) v.store(0);
  for (auto& v : in_flight) v.store(0);
  std::atomic<int>  done{0};
  std::atomic<long> double_take{0};

  auto finished = [&]() { return done.load(std::memory_order_relaxed) >= lanes || g_abort.load(); };
  // a "warp" = 32 tickets polled round-robin by one thread
  auto consumer = [&](Ring& from, int role) {
    std::vector<Ticket> t(32);
    unsigned            rng = 12345u + role * 977u;
    while (!finished()) {
      bool any = false;
      for (auto& tk : t) {
        int lane;
        if (!ring_take(from, tk, lane)) continue;
        any = true;
        if (in_flight[lane].fetch_add(1) != 0) double_take++;
        rng = rng * 1664525u + 1013904223u;
        int h = visits[lane].fetch_add(1) + 1;
        in_flight[lane].fetch_sub(1);
        if (role == 0) {  // traversal: hit -> shade
          ring_push(shade, lane);
        } else if (role == 1) {  // shade: path ends / light pdf / next ray
          if (h >= hops) done.fetch_add(1);
          else if ((rng >> 16) & 1) ring_push(lpdf, lane);
          else ring_push(ext, lane);
        } else {  // light pdf: -> next ray
          ring_push(ext, lane);
        }
      }
      if (!any) std::this_thread::yield();
    }
  };
  std::vector<std::thread> pool;
  for (int k = 0; k < threads; k++) {
    pool.emplace_back(consumer, std::ref(ext), 0);
    pool.emplace_back(consumer, std::ref(shade), 1);
    pool.emplace_back(consumer, std::ref(lpdf), 2);
  }
  for (int l = 0; l < lanes; l++) ring_push(ext, l);  // seed
  auto t0 = std::chrono::steady_clock::now();
  while (!finished()) {
    std::this_thread::sleep_for(std::chrono::milliseconds(5));
    if (std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() > 60) g_abort.store(2);
  }
  for (auto& th : pool) th.join();
  long bad = 0;
  for (int l = 0; l < lanes; l++) {
    // a lane ends in the shade role at its first shade visit with h >= hops; every visit is counted once
    if (visits[l].load() < hops) bad++;
  }
  printf("lanes %d hops %d cap %u: done %d abort %d double_take %ld short %ld\n", lanes, hops, cap, done.load(),
      g_abort.load(), double_take.load(), bad);
  return (g_abort.load() || double_take.load() || bad || done.load() != lanes) ? 1 : 0;
}
