// tests/kat/queue_stress.cpp — ThreadSanitizer stress of urf_queue.cpp (built with -fsanitize=thread, no CUDA):
// P producers and one consumer around a stand-in batch function; both policies; exits 0 when every accepted scan was
// delivered exactly once with the right payload and per-producer order, and TSAN reported nothing (TSAN makes the exit
// code non-zero on a report).
// usage: queue_stress <producers> <scans per producer> <slots> <max_batch> <policy>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>
#include "../../include/urf.h"

// stand-ins for the CUDA side of liburf_b200.so (never reached: the queue is created with urf_queue_create_with)
extern "C" void* urf_pinned_alloc(size_t) { return nullptr; }
extern "C" void urf_pinned_free(void*) {}
extern "C" int urf_process_batch(urf_ctx*, const float* const*, const int*, int, urf_result*) { return URF_ERR_NO_DEVICE; }

static std::atomic<int> g_batches{0};
static int fake(void*, const float* const* xyzi, const int* n, int batch, urf_result* outs) {
  g_batches++;
  for (int j = 0; j < batch; j++) {
    for (int i = 0; i < n[j]; i++) outs[j].label[i] = (int)xyzi[j][4 * i] + 7;
    outs[j].status = URF_OK; outs[j].n_in = n[j];
  }
  if ((g_batches.load() & 7) == 0) std::this_thread::yield();
  return URF_OK;
}

int main(int argc, char** argv) {
  const int P = argc > 1 ? atoi(argv[1]) : 4, K = argc > 2 ? atoi(argv[2]) : 2000, slots = argc > 3 ? atoi(argv[3]) : 6,
            mb = argc > 4 ? atoi(argv[4]) : 4, policy = argc > 5 ? atoi(argv[5]) : URF_QUEUE_BLOCK;
  const int N = 24;
  urf_queue* q = nullptr;
  if (urf_queue_create_with(&q, fake, nullptr, N, slots, mb, policy) != URF_OK) return 2;
  std::atomic<long> accepted{0};
  std::vector<std::thread> prod;
  for (int p = 0; p < P; p++) prod.emplace_back([&, p] {
    std::vector<float> pts(4 * N);
    for (int k = 0; k < K; k++) {
      const int n = 1 + (k + p) % N;
      for (int i = 0; i < n; i++) pts[4 * i] = (float)(k % 1000 + i);
      const int rc = urf_queue_submit(q, pts.data(), n, ((uint64_t)p << 32) | (uint64_t)k, -1);
      if (rc != URF_OK) { fprintf(stderr, "submit rc=%d\n", rc); exit(3); }
      accepted++;
    }
  });
  long delivered = 0, bad = 0;
  std::vector<long> last(P, -1);
  std::thread cons([&] {
    std::vector<int32_t> lab(N);
    for (;;) {
      urf_result r{}; r.label = lab.data();
      uint64_t tag = 0;
      const int rc = urf_queue_next(q, &tag, &r, -1);
      if (rc == URF_ERR_CLOSED) break;
      if (rc != URF_OK) { bad++; continue; }
      const int p = (int)(tag >> 32); const long k = (long)(tag & 0xffffffffu);
      if (k <= last[p]) bad++;                               // per-producer order (drops may leave gaps)
      last[p] = k;
      const int n = 1 + (int)((k + p) % N);
      if (r.n_in != n) bad++;
      for (int i = 0; i < n; i++) if (lab[i] != (int)(k % 1000 + i) + 7) { bad++; break; }
      delivered++;
    }
  });
  for (auto& t : prod) t.join();
  urf_queue_close(q);
  cons.join();
  urf_queue_stats st{};
  urf_queue_get_stats(q, &st);
  urf_queue_destroy(q);
  const bool ok = bad == 0 && st.submitted == (uint64_t)accepted.load() && st.processed + st.dropped == st.submitted &&
                  st.delivered == (uint64_t)delivered && st.delivered == st.processed && (policy == URF_QUEUE_DROP_OLDEST || st.dropped == 0);
  printf("producers=%d scans=%ld delivered=%ld dropped=%llu batches=%llu largest_batch=%d bad=%ld %s\n", P, accepted.load(), delivered,
         (unsigned long long)st.dropped, (unsigned long long)st.batches, st.largest_batch, bad, ok ? "OK" : "FAIL");
  return ok ? 0 : 1;
}
