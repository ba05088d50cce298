// urf_queue.cpp — streaming ingest in front of urf_process_batch (include/urf.h). Host code only.
//
// Replaces the reference's depth-1 subscriber (`nh->subscribe(params::topicName, 1, &Detector::filtered, this)`,
// lidar_segmentation.cpp:53): scans that arrive while a scan is being processed are staged instead of dropped, and the
// worker hands everything that is pending to one batched call.
//
// Slot life cycle (all transitions under one mutex):
//   FREE -> FILLING (producer copies the scan, lock released) -> PENDING -> RUNNING (worker) -> DONE -> FREE (consumer)
//   PENDING -> FREE when URF_QUEUE_DROP_OLDEST needs room (the scan is counted as dropped, never delivered).
// Results are delivered in submission order: every accepted scan gets a sequence number when it becomes PENDING and the
// consumer waits for the smallest live one.
#include <chrono>
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

#include "../../include/urf.h"

namespace {
enum SlotState { FREE = 0, FILLING, PENDING, RUNNING, DONE, VIEWED };   // VIEWED: delivered, its labels lent to the consumer
struct Slot {
  SlotState state = FREE;
  uint64_t seq = 0, tag = 0;
  int n = 0, rc = URF_OK;
  float* in = nullptr;       // max_points * bytes_per_point bytes (pinned for the real queue)
  const float* ext = nullptr;   // urf_queue_submit_ref: the caller's buffer is used in place (no copy)
  int32_t* label = nullptr;  // max_points
  urf_result res{};
};
}  // namespace

struct urf_queue {
  urf_queue_process_fn fn = nullptr;
  void* user = nullptr;
  bool pinned = false;
  int max_points = 0, max_batch = 1, policy = URF_QUEUE_BLOCK;
  // record format of the scans: step == 0: (x, y, z, intensity) float4 points; step > 0: raw PointCloud2 records of `step`
  // bytes (urf_queue_create_cloud2), handed to urf_process_cloud2_batch and unpacked on the device
  int step = 0, ox = 0, oy = 4, oz = 8, oi = -1;
  size_t bytes_per_point = 16;
  std::vector<Slot> slots;
  std::mutex mu;
  std::condition_variable cv_free, cv_pending, cv_done;
  uint64_t next_seq = 1;       // sequence number of the next accepted scan
  int viewed = -1;             // slot lent out by urf_queue_next_view, given back on the consumer's next call
  bool closed = false;
  urf_queue_stats st{};
  std::thread worker;
};

namespace {

int real_process(void* user, const float* const* xyzi, const int* n, int batch, urf_result* outs) {
  return urf_process_batch(static_cast<urf_ctx*>(user), xyzi, n, batch, outs);
}

template <class Pred>
bool wait_for(std::condition_variable& cv, std::unique_lock<std::mutex>& lk, int timeout_ms, Pred pred) {
  if (timeout_ms < 0) { cv.wait(lk, pred); return true; }
  return cv.wait_for(lk, std::chrono::milliseconds(timeout_ms), pred);
}

void worker_loop(urf_queue* q) {
  std::vector<int> idx;
  std::vector<const float*> ptrs;
  std::vector<int> ns;
  std::vector<urf_result> outs;
  for (;;) {
    idx.clear();
    {
      std::unique_lock<std::mutex> lk(q->mu);
      q->cv_pending.wait(lk, [&] {
        if (q->closed) return true;
        for (const Slot& s : q->slots) if (s.state == PENDING) return true;
        return false;
      });
      // everything pending, oldest first, up to max_batch
      for (;;) {
        int best = -1;
        for (int i = 0; i < (int)q->slots.size(); i++) {
          const Slot& s = q->slots[i];
          if (s.state == PENDING && (best < 0 || s.seq < q->slots[best].seq)) best = i;
        }
        if (best < 0 || (int)idx.size() >= q->max_batch) break;
        q->slots[best].state = RUNNING;
        idx.push_back(best);
      }
      if (idx.empty()) { if (q->closed) return; continue; }
      q->st.batches++;
      if ((int)idx.size() > q->st.largest_batch) q->st.largest_batch = (int)idx.size();
    }
    const int B = (int)idx.size();
    ptrs.resize(B); ns.resize(B); outs.assign(B, urf_result{});
    for (int j = 0; j < B; j++) {
      Slot& s = q->slots[idx[j]];
      ptrs[j] = s.ext ? s.ext : s.in; ns[j] = s.n;
      outs[j].label = s.label;
    }
    const int rc = q->step == 0 ? q->fn(q->user, ptrs.data(), ns.data(), B, outs.data())
                                : urf_process_cloud2_batch(static_cast<urf_ctx*>(q->user), reinterpret_cast<const void* const*>(ptrs.data()), ns.data(), B,
                                                           q->step, q->ox, q->oy, q->oz, q->oi, outs.data(), nullptr);
    {
      std::lock_guard<std::mutex> lk(q->mu);
      for (int j = 0; j < B; j++) {
        Slot& s = q->slots[idx[j]];
        s.res = outs[j]; s.rc = rc; s.state = DONE;
        q->st.processed++;
      }
    }
    q->cv_done.notify_all();
  }
}

int create_common(urf_queue** out, urf_queue_process_fn fn, void* user, bool pinned, int max_points, int slots, int max_batch, int policy,
                  int step = 0, int ox = 0, int oy = 4, int oz = 8, int oi = -1) {
  if (!out || !fn || max_points < 1 || slots < 1 || max_batch < 1 || (policy != URF_QUEUE_BLOCK && policy != URF_QUEUE_DROP_OLDEST))
    return URF_ERR_INVALID;
  urf_queue* q = new urf_queue;
  q->fn = fn; q->user = user; q->pinned = pinned; q->max_points = max_points; q->max_batch = max_batch; q->policy = policy;
  q->step = step; q->ox = ox; q->oy = oy; q->oz = oz; q->oi = oi;
  q->bytes_per_point = step > 0 ? (size_t)step : 16;
  q->slots.resize(slots);
  for (Slot& s : q->slots) {
    const size_t in_bytes = q->bytes_per_point * (size_t)max_points, lab_bytes = sizeof(int32_t) * (size_t)max_points;
    s.in = static_cast<float*>(pinned ? urf_pinned_alloc(in_bytes) : std::malloc(in_bytes));
    s.label = static_cast<int32_t*>(pinned ? urf_pinned_alloc(lab_bytes) : std::malloc(lab_bytes));
    if (!s.in || !s.label) {
      for (Slot& t : q->slots) {
        if (pinned) { urf_pinned_free(t.in); urf_pinned_free(t.label); } else { std::free(t.in); std::free(t.label); }
      }
      delete q;
      return URF_ERR_NOMEM;
    }
  }
  q->worker = std::thread(worker_loop, q);
  *out = q;
  return URF_OK;
}

}  // namespace

extern "C" {

int urf_queue_create(urf_queue** out, urf_ctx* ctx, int max_points, int slots, int max_batch, int policy) {
  if (!ctx) return URF_ERR_INVALID;
  return create_common(out, real_process, ctx, true, max_points, slots, max_batch, policy);
}

int urf_queue_create_cloud2(urf_queue** out, urf_ctx* ctx, int max_points, int slots, int max_batch, int policy, int point_step, int off_x,
                            int off_y, int off_z, int off_intensity) {
  if (!ctx || point_step < 12 || point_step > URF_MAX_POINT_STEP) return URF_ERR_INVALID;
  for (int o : {off_x, off_y, off_z}) if (o < 0 || o + 4 > point_step) return URF_ERR_INVALID;
  if (off_intensity >= 0 && off_intensity + 4 > point_step) return URF_ERR_INVALID;
  return create_common(out, real_process, ctx, true, max_points, slots, max_batch, policy, point_step, off_x, off_y, off_z, off_intensity);
}

int urf_queue_create_with(urf_queue** out, urf_queue_process_fn fn, void* user, int max_points, int slots, int max_batch, int policy) {
  return create_common(out, fn, user, false, max_points, slots, max_batch, policy);
}

namespace {
int submit_common(urf_queue* q, const void* data, int n, uint64_t tag, int timeout_ms, bool by_reference) {
  if (!q || n < 0 || (n > 0 && !data)) return URF_ERR_INVALID;
  if (n > q->max_points) return URF_ERR_CAPACITY;
  int slot = -1;
  {
    std::unique_lock<std::mutex> lk(q->mu);
    auto find = [&] {
      if (q->closed) return true;
      for (int i = 0; i < (int)q->slots.size(); i++) if (q->slots[i].state == FREE) { slot = i; return true; }
      if (q->policy == URF_QUEUE_DROP_OLDEST) {           // lidar_segmentation.cpp:53: the subscriber keeps only the newest scan
        int best = -1;
        for (int i = 0; i < (int)q->slots.size(); i++) {
          const Slot& s = q->slots[i];
          if (s.state == PENDING && (best < 0 || s.seq < q->slots[best].seq)) best = i;
        }
        if (best >= 0) { q->st.dropped++; slot = best; return true; }
      }
      return false;
    };
    if (!wait_for(q->cv_free, lk, timeout_ms, find)) return URF_ERR_TIMEOUT;
    if (q->closed) return URF_ERR_CLOSED;
    q->slots[slot].state = FILLING;                       // a dropped scan's sequence number simply never reaches DONE
  }
  Slot& s = q->slots[slot];
  bool closed_late = false;
  if (by_reference) s.ext = static_cast<const float*>(data);
  else { s.ext = nullptr; if (n > 0) std::memcpy(s.in, data, q->bytes_per_point * (size_t)n); }
  {
    std::lock_guard<std::mutex> lk(q->mu);
    if (q->closed) {                                      // closed while copying: the worker may already be gone
      s.state = FREE;
      closed_late = true;
    } else {
      s.n = n; s.tag = tag; s.rc = URF_OK;
      s.seq = q->next_seq++;
      s.state = PENDING;
      q->st.submitted++;
    }
  }
  if (closed_late) {
    // a consumer whose last look at the slots still saw this one FILLING must get to see the drained state: close()'s
    // own notify may have come before that look
    q->cv_done.notify_all();
    q->cv_free.notify_one();
    return URF_ERR_CLOSED;
  }
  q->cv_pending.notify_one();
  q->cv_done.notify_all();                                // a consumer waiting on a dropped sequence number re-evaluates
  return URF_OK;
}
}  // namespace

int urf_queue_submit(urf_queue* q, const float* xyzi, int n, uint64_t tag, int timeout_ms) {
  if (q && q->step != 0) return URF_ERR_INVALID;           // a record queue takes urf_queue_submit_cloud2
  return submit_common(q, xyzi, n, tag, timeout_ms, false);
}

int urf_queue_submit_ref(urf_queue* q, const float* xyzi, int n, uint64_t tag, int timeout_ms) {
  if (q && q->step != 0) return URF_ERR_INVALID;
  return submit_common(q, xyzi, n, tag, timeout_ms, true);
}

int urf_queue_submit_cloud2(urf_queue* q, const void* data, int n_points, uint64_t tag, int timeout_ms) {
  if (q && q->step == 0) return URF_ERR_INVALID;
  return submit_common(q, data, n_points, tag, timeout_ms, false);
}

namespace {
int next_common(urf_queue* q, uint64_t* tag, urf_result* out, const int32_t** label_view, int timeout_ms);
}

int urf_queue_next(urf_queue* q, uint64_t* tag, urf_result* out, int timeout_ms) { return next_common(q, tag, out, nullptr, timeout_ms); }

int urf_queue_next_view(urf_queue* q, uint64_t* tag, urf_result* out, const int32_t** label_view, int timeout_ms) {
  if (!label_view) return URF_ERR_INVALID;
  return next_common(q, tag, out, label_view, timeout_ms);
}

namespace {
// label_view != NULL: no copy — *label_view points at the labels inside the queue's staging slot, which stays reserved
// (not reusable by producers) until this consumer's next urf_queue_next* call on the queue.
int next_common(urf_queue* q, uint64_t* tag, urf_result* out, const int32_t** label_view, int timeout_ms) {
  if (!q || !out) return URF_ERR_INVALID;
  std::unique_lock<std::mutex> lk(q->mu);
  if (q->viewed >= 0) {                                   // the slot lent out by the previous view call comes back now
    q->slots[q->viewed].state = FREE;
    q->viewed = -1;
    q->cv_free.notify_one();
  }
  int slot = -1;
  bool drained = false;
  auto ready = [&] {
    // the oldest live scan: smallest sequence number among PENDING / RUNNING / DONE slots
    int best = -1;
    bool filling = false;
    for (int i = 0; i < (int)q->slots.size(); i++) {
      const Slot& s = q->slots[i];
      if (s.state == FILLING) filling = true;
      if ((s.state == PENDING || s.state == RUNNING || s.state == DONE) && (best < 0 || s.seq < q->slots[best].seq)) best = i;
    }
    if (best >= 0 && q->slots[best].state == DONE) { slot = best; return true; }
    if (best < 0 && !filling && q->closed) { drained = true; return true; }
    return false;
  };
  if (!wait_for(q->cv_done, lk, timeout_ms, ready)) return URF_ERR_TIMEOUT;
  if (drained) return URF_ERR_CLOSED;
  Slot& s = q->slots[slot];
  int32_t* user_label = out->label;
  const int rc = s.rc;
  *out = s.res;
  out->label = user_label; out->ring = nullptr; out->order = nullptr; out->ring_start = nullptr;
  if (tag) *tag = s.tag;
  q->st.delivered++;
  if (label_view) {                                       // lend the slot: DONE slots are invisible to producers and the worker
    *label_view = rc == URF_OK ? s.label : nullptr;
    s.state = VIEWED;
    q->viewed = slot;
    return rc;
  }
  const int n = s.n;
  const int32_t* src = s.label;
  s.state = VIEWED;                                       // ours: invisible to producers, the worker and other consumers
  lk.unlock();                                            // the copy runs outside the lock
  if (user_label && rc == URF_OK && n > 0) std::memcpy(user_label, src, sizeof(int32_t) * (size_t)n);
  lk.lock();
  s.state = FREE;
  lk.unlock();
  q->cv_free.notify_one();
  return rc;
}
}  // namespace

void urf_queue_release_view(urf_queue* q) {
  if (!q) return;
  std::lock_guard<std::mutex> lk(q->mu);
  if (q->viewed >= 0) { q->slots[q->viewed].state = FREE; q->viewed = -1; q->cv_free.notify_one(); }
}

int urf_queue_get_stats(urf_queue* q, urf_queue_stats* st) {
  if (!q || !st) return URF_ERR_INVALID;
  std::lock_guard<std::mutex> lk(q->mu);
  *st = q->st;
  st->pending = 0;
  for (const Slot& s : q->slots) if (s.state == PENDING || s.state == RUNNING || s.state == FILLING) st->pending++;
  return URF_OK;
}

void urf_queue_close(urf_queue* q) {
  if (!q) return;
  {
    std::lock_guard<std::mutex> lk(q->mu);
    q->closed = true;
  }
  q->cv_pending.notify_all();
  q->cv_free.notify_all();
  q->cv_done.notify_all();
}

void urf_queue_destroy(urf_queue* q) {
  if (!q) return;
  urf_queue_close(q);
  if (q->worker.joinable()) q->worker.join();             // the worker drains what is pending before it returns
  for (Slot& s : q->slots) {
    if (q->pinned) { urf_pinned_free(s.in); urf_pinned_free(s.label); } else { std::free(s.in); std::free(s.label); }
  }
  delete q;
}

}  // extern "C"
