// urf_mq.cpp — ONE ingest stream over SEVERAL GPUs (include/urf.h urf_mq, BASELINE config 4: a continuous scan stream
// sharded across the B200s of one box). Host code only.
//
// The reference is a single subscriber with queue depth 1 (`nh->subscribe(params::topicName, 1, &Detector::filtered, this)`,
// lidar_segmentation.cpp:53); the demo graph feeds four LiDAR topics (config/demo1.rviz:91,121,151,181). urf_mq keeps one
// submit/next interface in front of N devices: every device owns a context and a streaming queue (urf_queue: pinned
// staging slots + one worker thread that batches whatever is pending through urf_process_batch); a scan goes to the device
// with the fewest scans in flight (ties: round-robin), and results come back in the order the submissions completed.
// Scans are independent units, so there is nothing to exchange between devices (no collective on this path).
//
// Ordering: urf_queue delivers each device's scans in the order their submit calls completed, so the global order only
// has to remember WHICH device holds the next scan: a FIFO of device indices, appended after a device accepted a scan
// (under that device's submit mutex, so that the k-th entry naming a device is the k-th scan its queue accepted).
#include <cstring>
#include <deque>
#include <mutex>
#include <condition_variable>
#include <chrono>
#include <vector>

#include "../../include/urf.h"

struct urf_mq {
  struct Dev {
    int device = -1;
    urf_ctx* ctx = nullptr;          // owned (NULL with the test hook)
    urf_queue* q = nullptr;
    uint64_t submitted = 0, delivered = 0;
    int inflight = 0;                // accepted or being copied, not yet delivered
    std::mutex submit_mu;            // held from the device queue's submit to the append to `order`: one device's entries
                                     // enter `order` in the order its queue accepted them (copies to DIFFERENT devices overlap)
  };
  std::deque<Dev> dev;               // deque: Dev holds a mutex and must not move
  std::mutex mu;
  std::condition_variable cv;        // a device index was appended / the mq was closed
  std::deque<int> order;             // device of the next scans to deliver, oldest first
  int rr = 0;                        // round-robin cursor for ties
  int submitting = 0;                // submit calls between device choice and order append
  bool closed = false;
  int view_dev = -1;                 // consumer side only: device whose queue has lent out a label view
};

namespace {

int pick_device(urf_mq* m) {         // fewest scans in flight; ties go round-robin so an idle box is loaded evenly
  const int n = (int)m->dev.size();
  int best = -1;
  for (int j = 0; j < n; j++) {
    const int d = (m->rr + j) % n;
    if (best < 0 || m->dev[d].inflight < m->dev[best].inflight) best = d;
  }
  m->rr = (best + 1) % n;
  return best;
}

int submit_common(urf_mq* m, const float* xyzi, int n, uint64_t tag, int timeout_ms, bool by_reference) {
  if (!m) return URF_ERR_INVALID;
  int d;
  {
    std::lock_guard<std::mutex> lk(m->mu);
    if (m->closed) return URF_ERR_CLOSED;
    d = pick_device(m);
    m->dev[d].inflight++;
    m->submitting++;
  }
  // outside the mq lock: the copy into the device's pinned slot (or the wait for a free slot) runs in parallel for
  // producers that were dealt different devices
  std::lock_guard<std::mutex> dev_lk(m->dev[d].submit_mu);
  const int rc = by_reference ? urf_queue_submit_ref(m->dev[d].q, xyzi, n, tag, timeout_ms) : urf_queue_submit(m->dev[d].q, xyzi, n, tag, timeout_ms);
  {
    std::lock_guard<std::mutex> lk(m->mu);
    m->submitting--;
    if (rc == URF_OK) { m->order.push_back(d); m->dev[d].submitted++; }
    else m->dev[d].inflight--;
  }
  m->cv.notify_all();
  return rc;
}

}  // namespace

extern "C" {

int urf_mq_create(urf_mq** out, const int* devices, int n_devices, int max_points, int slots_per_device, int max_batch,
                  const urf_params* params) {
  if (!out || !devices || n_devices < 1 || max_points < 1 || slots_per_device < 1 || max_batch < 1) return URF_ERR_INVALID;
  *out = nullptr;
  urf_mq* m = new urf_mq;
  for (int j = 0; j < n_devices; j++) m->dev.emplace_back();
  int rc = URF_OK;
  for (int j = 0; j < n_devices && rc == URF_OK; j++) {
    urf_mq::Dev& d = m->dev[j];
    d.device = devices[j];
    rc = urf_create(&d.ctx, d.device, max_points, max_batch);
    if (rc == URF_OK && params) rc = urf_set_params(d.ctx, params);
    if (rc == URF_OK) rc = urf_queue_create(&d.q, d.ctx, max_points, slots_per_device, max_batch, URF_QUEUE_BLOCK);
  }
  if (rc != URF_OK) { urf_mq_destroy(m); return rc; }
  *out = m;
  return URF_OK;
}

int urf_mq_create_with(urf_mq** out, urf_queue_process_fn fn, void* const* users, int n_devices, int max_points, int slots_per_device,
                       int max_batch) {
  if (!out || !fn || n_devices < 1) return URF_ERR_INVALID;
  *out = nullptr;
  urf_mq* m = new urf_mq;
  for (int j = 0; j < n_devices; j++) m->dev.emplace_back();
  for (int j = 0; j < n_devices; j++) {
    m->dev[j].device = j;
    const int rc = urf_queue_create_with(&m->dev[j].q, fn, users ? users[j] : nullptr, max_points, slots_per_device, max_batch, URF_QUEUE_BLOCK);
    if (rc != URF_OK) { urf_mq_destroy(m); return rc; }
  }
  *out = m;
  return URF_OK;
}

int urf_mq_set_params(urf_mq* m, const urf_params* p) {
  if (!m || !p) return URF_ERR_INVALID;
  // like the reference's paramsCallback between two scan callbacks (single spin thread, src/main.cpp:54): the caller
  // reconfigures between scans — everything submitted must have been collected
  {
    std::lock_guard<std::mutex> lk(m->mu);
    if (!m->order.empty() || m->submitting) return URF_ERR_INVALID;
  }
  for (urf_mq::Dev& d : m->dev)
    if (d.ctx) { const int rc = urf_set_params(d.ctx, p); if (rc != URF_OK) return rc; }
  return URF_OK;
}

int urf_mq_submit(urf_mq* m, const float* xyzi, int n, uint64_t tag, int timeout_ms) { return submit_common(m, xyzi, n, tag, timeout_ms, false); }
int urf_mq_submit_ref(urf_mq* m, const float* xyzi, int n, uint64_t tag, int timeout_ms) { return submit_common(m, xyzi, n, tag, timeout_ms, true); }

namespace {
int mq_next_common(urf_mq* m, uint64_t* tag, urf_result* out, const int32_t** label_view, int timeout_ms);
}
int urf_mq_next(urf_mq* m, uint64_t* tag, urf_result* out, int timeout_ms) { return mq_next_common(m, tag, out, nullptr, timeout_ms); }
int urf_mq_next_view(urf_mq* m, uint64_t* tag, urf_result* out, const int32_t** label_view, int timeout_ms) {
  if (!label_view) return URF_ERR_INVALID;
  return mq_next_common(m, tag, out, label_view, timeout_ms);
}
namespace {
int mq_next_common(urf_mq* m, uint64_t* tag, urf_result* out, const int32_t** label_view, int timeout_ms) {
  if (!m || !out) return URF_ERR_INVALID;
  int d;
  {
    std::unique_lock<std::mutex> lk(m->mu);
    auto ready = [&] { return !m->order.empty() || (m->closed && m->submitting == 0); };
    if (timeout_ms < 0) m->cv.wait(lk, ready);
    else if (!m->cv.wait_for(lk, std::chrono::milliseconds(timeout_ms), ready)) return URF_ERR_TIMEOUT;
    if (m->order.empty()) return URF_ERR_CLOSED;          // closed and drained
    d = m->order.front();
  }
  // a view lent by the previous call belongs to one device's queue: give it back before another device's queue lends one
  if (m->view_dev >= 0 && (m->view_dev != d || !label_view)) { urf_queue_release_view(m->dev[m->view_dev].q); m->view_dev = -1; }
  const int rc = label_view ? urf_queue_next_view(m->dev[d].q, tag, out, label_view, timeout_ms) : urf_queue_next(m->dev[d].q, tag, out, timeout_ms);
  if (rc == URF_ERR_TIMEOUT) return rc;                   // still the oldest scan: the entry stays at the front
  if (label_view) m->view_dev = d;
  {
    std::lock_guard<std::mutex> lk(m->mu);
    m->order.pop_front();
    m->dev[d].inflight--;
    m->dev[d].delivered++;
  }
  return rc;
}
}  // namespace

int urf_mq_get_stats(urf_mq* m, urf_mq_stats* st) {
  if (!m || !st) return URF_ERR_INVALID;
  std::memset(st, 0, sizeof(*st));
  std::lock_guard<std::mutex> lk(m->mu);
  st->n_devices = (int32_t)m->dev.size();
  for (size_t j = 0; j < m->dev.size() && j < URF_MQ_MAX_DEVICES; j++) {
    st->submitted[j] = m->dev[j].submitted; st->delivered[j] = m->dev[j].delivered;
    urf_queue_stats qs{};
    if (m->dev[j].q && urf_queue_get_stats(m->dev[j].q, &qs) == URF_OK) { st->batches[j] = qs.batches; st->largest_batch[j] = qs.largest_batch; }
  }
  st->pending = (int32_t)m->order.size();
  return URF_OK;
}

void urf_mq_close(urf_mq* m) {
  if (!m) return;
  {
    std::lock_guard<std::mutex> lk(m->mu);
    m->closed = true;
  }
  for (urf_mq::Dev& d : m->dev) if (d.q) urf_queue_close(d.q);
  m->cv.notify_all();
}

void urf_mq_destroy(urf_mq* m) {
  if (!m) return;
  urf_mq_close(m);
  for (urf_mq::Dev& d : m->dev) {
    if (d.q) urf_queue_destroy(d.q);
    if (d.ctx) urf_destroy(d.ctx);
  }
  delete m;
}

}  // extern "C"
