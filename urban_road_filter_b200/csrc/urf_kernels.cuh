// urf_kernels.cuh — sm_100a kernels of the per-scan road/curb classification path.
//
// Every kernel takes the batch index from blockIdx.y (or blockIdx.x for one-CTA-per-scan kernels): a launch covers a
// whole batch of scans laid out back to back with `S` points of stride. All arithmetic that decides a label uses
// round-to-nearest intrinsics in exactly the operation order of the reference (file:line cited per block; paths are
// relative to the reference repo), so results are bit-identical to the x86-64 -O2 build of the reference.
//
// Pipeline (DESIGN.md has the picture):
//   k_reset -> k_points -> k_register -> k_assign -> [k_register_exact -> k_assign(redo)] -> k_scan_offsets -> k_scatter
//   -> k_star_sort -> k_star_scan -> k_ring_detect -> k_tables -> k_label -> k_cutkey -> k_dmax -> k_best -> k_verts
//   [-> k_sort_rings when the emission order is requested]
#pragma once
#include "urf_device.cuh"
#include "urf_logic.cuh"

namespace urf {

__constant__ float c_beam_d[kSectKeys];
__constant__ float c_beam_o[kSectKeys];
__constant__ unsigned char c_beam_yx[kSectKeys];

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }

// CTA-wide bitonic sort of npad (power of two) keys in shared or global memory; all threads must call.
template <class T>
__device__ void cta_bitonic(T* a, int npad) {
  for (int k = 2; k <= npad; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = threadIdx.x; t < (npad >> 1); t += blockDim.x) {
        int i = 2 * t - (t & (j - 1));
        int l = i + j;
        bool up = (i & k) == 0;
        T x = a[i], y = a[l];
        if ((x > y) == up) { a[i] = y; a[l] = x; }
      }
      __syncthreads();
    }
  }
}

__device__ __forceinline__ int next_pow2(int n) { int p = 1; while (p < n) p <<= 1; return p; }

// largest k in [0, n_rings) with ring_start[k] <= p (ring_start ascending, ring_start[0] == 0)
__device__ __forceinline__ int ring_of(const int* ring_start, int n_rings, int p) {
  int lo = 0, hi = n_rings;      // invariant: ring_start[lo] <= p < ring_start[hi]
  while (hi - lo > 1) {
    int mid = (lo + hi) >> 1;
    if (ring_start[mid] <= p) lo = mid; else hi = mid;
  }
  return lo;
}

// ---------------------------------------------------------------------------------------------------------------------
// k_reset: per-call initialisation of the per-scan tables.
__global__ void k_reset(DevBuffers buf, DevParams prm) {
  const int b = blockIdx.y;
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  const int nth = gridDim.x * blockDim.x;
  ScanOut& o = buf.out[b];
  ScanTab& t = buf.tab[b];
  if (tid == 0) {
    o.n_in = buf.n[b]; o.n_roi = 0; o.n_rings = 0; o.n_order = 0; o.n_road = 0; o.n_curb = 0; o.n_vert = 0; o.flags = 0;
  }
  for (int i = tid; i <= kRingKeys; i += nth) o.ring_start[i] = 0;
  for (int i = tid; i < kRingKeys; i += nth) { t.maxdist[i] = 0u; t.angle[i] = 0.f; t.regidx[i] = 0x7fffffff; t.regorder[i] = 0x7fffffff; }
  for (int i = tid; i < kDegBins; i += nth) {
    t.cut[i] = 0x7fffffff; t.cutkey[i] = ~0ull; t.dmax[i] = 0u; t.best[i] = ~0ull;
  }
  unsigned* fi = buf.firstidx + (size_t)b * (kElevBins + 1);
  for (int i = tid; i <= kElevBins; i += nth) fi[i] = 0xffffffffu;
  const size_t nb = (size_t)prm.channels * kDegBins;
  unsigned* cmin = buf.cmin + (size_t)b * nb;
  unsigned* cmax = buf.cmax + (size_t)b * nb;
  for (size_t i = tid; i < nb; i += nth) { cmin[i] = 0x7f800000u; cmax[i] = 0u; }
}

// ---------------------------------------------------------------------------------------------------------------------
// k_points: ROI crop predicate + range + elevation angle per input point.
//   ROI: lidar_segmentation.cpp:106-113 (+ PCL ConditionalRemoval drops non-finite xyz)
//   d, alpha: lidar_segmentation.cpp:148-166
// Also records, per fine elevation bin, the first input index that falls into it (speculation input for k_register).
__global__ void k_points(DevBuffers buf, DevParams prm, int S) {
  const int b = blockIdx.y;
  const int n = buf.n[b];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  int keep = 0;
  if (i < n) {
    const size_t g = (size_t)b * S + i;
    const float4 p = __ldg(&buf.in[g]);
    keep = roi_keep(prm, p.x, p.y, p.z);
    float a = -1.0f;
    if (keep) {
      a = elev_alpha(p.x, p.y, p.z);
      int bin = (int)(a * (kElevBins / 180.0f));
      bin = bin < 0 ? 0 : (bin > kElevBins ? kElevBins : bin);
      unsigned* fi = buf.firstidx + (size_t)b * (kElevBins + 1) + bin;
      if (*(volatile unsigned*)fi > (unsigned)i) atomicMin(fi, (unsigned)i);
      if (a == 0.0f) atomicOr(&buf.out[b].flags, F_ZERO_ALPHA);
    }
    buf.alpha_v[g] = a;
    buf.mark[g] = 0;
  }
  const unsigned bal = __ballot_sync(0xffffffffu, keep);
  __shared__ int s_cnt;
  if (threadIdx.x == 0) s_cnt = 0;
  __syncthreads();
  if (lane_id() == 0 && bal) atomicAdd(&s_cnt, __popc(bal));
  __syncthreads();
  if (threadIdx.x == 0 && s_cnt) atomicAdd(&buf.out[b].n_roi, s_cnt);
}

// ---------------------------------------------------------------------------------------------------------------------
// Ring registration, lidar_segmentation.cpp:136-139,170-196: a point registers its elevation angle iff no VISIBLE
// registered angle lies within `interval` of it and fewer than `channels` angles are registered. "Visible": the
// reference's scan stops at the first angle[j] == 0, so once an angle of exactly 0 is registered, it and everything
// registered after it are never compared again.
//
// Exact path (any input): rounds of "find the first uncovered point after the last registrant" with a CTA-wide min.
__device__ void register_exact_cta(const float* __restrict__ alpha, int n, float interval, int channels, float* s_vis,
                                   float* s_reg, int* s_idx, int* s_red, int* out_m) {
  __shared__ int s_min;
  int m = 0, vis = 0, i_last = -1;
  bool frozen = false;
  while (m < channels) {
    int local = 0x7fffffff;
    for (int i = i_last + 1 + threadIdx.x; i < n; i += blockDim.x) {
      const float a = alpha[i];
      if (a < 0.0f) continue;
      bool cov = false;
      for (int t = 0; t < vis; t++) {
        if (fabsf(__fsub_rn(s_vis[t], a)) <= interval) { cov = true; break; }   // :179
      }
      if (!cov) { local = i; break; }
    }
    // CTA min
    for (int o = 16; o > 0; o >>= 1) local = min(local, __shfl_xor_sync(0xffffffffu, local, o));
    if (lane_id() == 0) s_red[threadIdx.x >> 5] = local;
    __syncthreads();
    if (threadIdx.x == 0) {
      int v = 0x7fffffff;
      for (int w = 0; w < (int)(blockDim.x >> 5); w++) v = min(v, s_red[w]);
      s_min = v;
    }
    __syncthreads();
    const int istar = s_min;
    __syncthreads();
    if (istar == 0x7fffffff) break;
    const float a = alpha[istar];
    if (threadIdx.x == 0) { s_reg[m] = a; s_idx[m] = istar; if (!frozen && a != 0.0f) s_vis[vis] = a; }
    if (!frozen) { if (a == 0.0f) frozen = true; else vis++; }
    m++;
    i_last = istar;
    __syncthreads();
  }
  *out_m = m;
}

// Sort the m registered (angle, regidx) pairs by angle (:205) and publish them. Angles are >= 0 so bits order them.
__device__ void publish_rings_cta(ScanTab& tab, ScanOut& out, const float* s_reg, const int* s_idx, int m,
                                  unsigned long long* s_keys) {
  for (int t = threadIdx.x; t < kRingKeys; t += blockDim.x)
    s_keys[t] = t < m ? (((unsigned long long)fbits(s_reg[t]) << 32) | (unsigned)s_idx[t]) : ~0ull;
  __syncthreads();
  cta_bitonic(s_keys, kRingKeys);
  for (int t = threadIdx.x; t < kRingKeys; t += blockDim.x) {
    if (t < m) {
      tab.angle[t] = bitsf((unsigned)(s_keys[t] >> 32));
      tab.regidx[t] = (int)(unsigned)s_keys[t];
      tab.regorder[t] = s_idx[t];
    } else { tab.angle[t] = 0.f; tab.regidx[t] = 0x7fffffff; tab.regorder[t] = 0x7fffffff; }
  }
  if (threadIdx.x == 0) out.n_rings = m;
}

// k_register: one CTA (256 threads) per scan. Fast path: the greedy registration is run over "candidates" only — the
// first point of every non-empty fine elevation bin, in input order. That is a speculation (a registrant need not be
// the first of its bin); k_assign verifies it against every point and k_register_exact repairs a failed speculation.
__global__ void __launch_bounds__(256) k_register(DevBuffers buf, DevParams prm, int S) {
  const int b = blockIdx.x;
  const int n = buf.n[b];
  ScanOut& out = buf.out[b];
  ScanTab& tab = buf.tab[b];
  const float* alpha = buf.alpha_v + (size_t)b * S;
  __shared__ unsigned s_cand[kMaxCand];
  __shared__ float s_calpha[kMaxCand];
  __shared__ float s_vis[kRingKeys];
  __shared__ float s_reg[kRingKeys];
  __shared__ int s_idx[kRingKeys];
  __shared__ unsigned long long s_keys[kRingKeys];
  __shared__ int s_red[8];
  __shared__ int s_cnt, s_m;
  if (out.n_roi < 30) {                      // lidar_segmentation.cpp:124-126: nothing happens for this scan
    if (threadIdx.x == 0) out.n_rings = 0;
    return;
  }
  bool exact = prm.force_exact || (out.flags & F_ZERO_ALPHA);
  if (!exact) {
    if (threadIdx.x == 0) s_cnt = 0;
    __syncthreads();
    const unsigned* fi = buf.firstidx + (size_t)b * (kElevBins + 1);
    for (int t = threadIdx.x; t <= kElevBins; t += blockDim.x) {
      const unsigned v = fi[t];
      if (v != 0xffffffffu) { int s = atomicAdd(&s_cnt, 1); if (s < kMaxCand) s_cand[s] = v; }
    }
    __syncthreads();
    const int cnt = s_cnt;
    if (cnt > kMaxCand) exact = true;        // uniform across the CTA
    else {
      const int npad = next_pow2(cnt < 2 ? 2 : cnt);
      for (int t = cnt + threadIdx.x; t < npad; t += blockDim.x) s_cand[t] = 0xffffffffu;
      __syncthreads();
      cta_bitonic(s_cand, npad);
      for (int t = threadIdx.x; t < cnt; t += blockDim.x) s_calpha[t] = alpha[s_cand[t]];
      __syncthreads();
      if (threadIdx.x < 32) {
        int m = 0;
        for (int c = 0; c < cnt && m < prm.channels; c++) {
          const float a = s_calpha[c];
          bool cov = false;
          for (int t = lane_id(); t < m; t += 32)
            if (fabsf(__fsub_rn(s_reg[t], a)) <= prm.interval) cov = true;          // :179
          cov = __any_sync(0xffffffffu, cov);
          if (!cov) {
            if (lane_id() == 0) { s_reg[m] = a; s_idx[m] = (int)s_cand[c]; }
            m++;
            __syncwarp();
          }
        }
        if (lane_id() == 0) s_m = m;
      }
      __syncthreads();
    }
  }
  if (exact) {
    int m;
    register_exact_cta(alpha, n, prm.interval, prm.channels, s_vis, s_reg, s_idx, s_red, &m);
    if (threadIdx.x == 0) { s_m = m; atomicOr(&out.flags, F_EXACT_REG); }
    __syncthreads();
  }
  publish_rings_cta(tab, out, s_reg, s_idx, s_m, s_keys);
}

// k_register_exact: repairs scans whose speculation failed verification in k_assign (F_SPEC_VIOLATION).
__global__ void __launch_bounds__(256) k_register_exact(DevBuffers buf, DevParams prm, int S) {
  const int b = blockIdx.x;
  ScanOut& out = buf.out[b];
  if (!(out.flags & F_SPEC_VIOLATION) || (out.flags & F_EXACT_REG)) return;
  __shared__ float s_vis[kRingKeys];
  __shared__ float s_reg[kRingKeys];
  __shared__ int s_idx[kRingKeys];
  __shared__ unsigned long long s_keys[kRingKeys];
  __shared__ int s_red[8];
  __shared__ int s_m;
  int m;
  register_exact_cta(buf.alpha_v + (size_t)b * S, buf.n[b], prm.interval, prm.channels, s_vis, s_reg, s_idx, s_red, &m);
  if (threadIdx.x == 0) s_m = m;
  __syncthreads();
  publish_rings_cta(buf.tab[b], out, s_reg, s_idx, s_m, s_keys);
  // F_EXACT_REG is set by the redo pass of k_assign (it must still see "violation and not yet exact" in every CTA)
}

// ---------------------------------------------------------------------------------------------------------------------
// k_assign: per input point — ring index (lidar_segmentation.cpp:226-233: first sorted angle within `interval`),
// star-shaped sector (star_shaped_search.cpp:164-173 + rectangular beam filter :73-107), default label, and the
// per-warp-chunk key histograms of the two stable partitions. redo=1 re-runs only for scans that were repaired.
__global__ void __launch_bounds__(kWarpsPerBlock * 32) k_assign(DevBuffers buf, DevParams prm, int S, int T, int redo) {
  const int b = blockIdx.y;
  ScanOut& out = buf.out[b];
  const int flags = out.flags;
  if (redo && (!(flags & F_SPEC_VIOLATION) || (flags & F_EXACT_REG))) return;
  const int n = buf.n[b];
  const int warp = threadIdx.x >> 5, lane = lane_id();
  const int chunk = blockIdx.x * kWarpsPerBlock + warp;
  __shared__ float s_angle[kRingKeys];
  __shared__ int s_regidx[kRingKeys];
  __shared__ unsigned s_cnt[kWarpsPerBlock][kKeys];
  const ScanTab& tab = buf.tab[b];
  const int R = out.n_rings;
  for (int t = threadIdx.x; t < kRingKeys; t += blockDim.x) { s_angle[t] = tab.angle[t]; s_regidx[t] = tab.regidx[t]; }
  for (int t = lane; t < kKeys; t += 32) s_cnt[warp][t] = 0;
  __syncthreads();
  if (chunk * kChunk >= n) return;             // whole warp; no block-level sync follows
  const bool live = out.n_roi >= 30;
  const bool verify = !redo && !(flags & F_EXACT_REG);
  unsigned* cnt = s_cnt[warp];
  bool violation = false;
  for (int it = 0; it < kChunk / 32; it++) {
    const int i = chunk * kChunk + it * 32 + lane;
    int ring = -1, sec = -1;
    if (i < n) {
      const size_t g = (size_t)b * S + i;
      const float a = buf.alpha_v[g];
      const bool kept = live && a >= 0.0f;
      if (kept) {
        int lo;
        ring = assign_ring(s_angle, R, a, prm.interval, &lo);
        if (verify && registration_violation(s_angle, s_regidx, tab.regorder, R, prm.channels, prm.interval, a, i, lo))
          violation = true;
        if (prm.star) {
          const float4 p = __ldg(&buf.in[g]);
          sec = star_sector(prm, p.x, p.y, c_beam_d, c_beam_o, c_beam_yx);
        }
      }
      buf.ringid[g] = (short)ring;
      buf.sect[g] = (short)sec;
      buf.label[g] = kept ? URF_LABEL_NONE : URF_LABEL_OUTSIDE;
    }
    unsigned peers = __match_any_sync(0xffffffffu, ring);
    if (ring >= 0 && lane == __ffs(peers) - 1) cnt[ring] += __popc(peers);
    peers = __match_any_sync(0xffffffffu, sec);
    if (sec >= 0 && lane == __ffs(peers) - 1) cnt[kRingKeys + sec] += __popc(peers);
    __syncwarp();
  }
  if (__any_sync(0xffffffffu, violation) && lane == 0) atomicOr(&out.flags, F_SPEC_VIOLATION);
  unsigned* row = buf.hist + ((size_t)b * T + chunk) * kKeys;
  for (int t = lane; t < kKeys; t += 32) row[t] = cnt[t];
}

// After the redo pass: mark repaired scans as exact (separate tiny kernel so every k_assign(redo) CTA saw the old flags).
__global__ void k_mark_exact(DevBuffers buf, int B) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < B && (buf.out[b].flags & F_SPEC_VIOLATION)) buf.out[b].flags |= F_EXACT_REG;
}

// ---------------------------------------------------------------------------------------------------------------------
// k_scan_offsets: one CTA (1024 threads) per scan turns hist[chunk][key] into exclusive scatter offsets
// (key-major bases + prefix over chunks) and publishes ring_start / sect_start.
__global__ void __launch_bounds__(1024) k_scan_offsets(DevBuffers buf, int T) {
  extern __shared__ unsigned s_part[];          // [32][kKeys] per-warp partial sums, then per-warp exclusive prefixes
  __shared__ unsigned s_base[kKeys];
  const int b = blockIdx.x;
  const int n = buf.n[b];
  const int rows = (n + kChunk - 1) / kChunk;
  const int warp = threadIdx.x >> 5, lane = lane_id();
  const int rpw = (rows + 31) / 32;
  const int r0 = min(rows, warp * rpw), r1 = min(rows, (warp + 1) * rpw);
  unsigned* hist = buf.hist + (size_t)b * T * kKeys;
  for (int key = lane; key < kKeys; key += 32) {
    unsigned s = 0;
    for (int r = r0; r < r1; r++) s += hist[(size_t)r * kKeys + key];
    s_part[warp * kKeys + key] = s;
  }
  __syncthreads();
  if (threadIdx.x < kKeys) {
    unsigned run = 0;
    for (int w = 0; w < 32; w++) { unsigned v = s_part[w * kKeys + threadIdx.x]; s_part[w * kKeys + threadIdx.x] = run; run += v; }
    s_base[threadIdx.x] = run;                  // total of this key
  }
  __syncthreads();
  if (threadIdx.x == 0) {                        // key bases: rings and sectors are separate address spaces
    unsigned run = 0;
    ScanOut& o = buf.out[b];
    for (int k = 0; k < kRingKeys; k++) { unsigned v = s_base[k]; s_base[k] = run; o.ring_start[k] = (int)run; run += v; }
    o.ring_start[kRingKeys] = (int)run;
    o.n_order = (int)run;
    ScanTab& t = buf.tab[b];
    run = 0;
    for (int k = 0; k < kSectKeys; k++) { unsigned v = s_base[kRingKeys + k]; s_base[kRingKeys + k] = run; t.sect_start[k] = (int)run; run += v; }
    t.sect_start[kSectKeys] = (int)run;
  }
  __syncthreads();
  for (int key = lane; key < kKeys; key += 32) {
    unsigned run = s_base[key] + s_part[warp * kKeys + key];
    for (int r = r0; r < r1; r++) {
      unsigned* p = &hist[(size_t)r * kKeys + key];
      unsigned v = *p; *p = run; run += v;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// k_scatter: stable scatter of every point into its ring bucket (input order inside a ring, lidar_segmentation.cpp:221-
// 277) and into its star sector (push_back order, star_shaped_search.cpp:173).
__global__ void __launch_bounds__(kWarpsPerBlock * 32) k_scatter(DevBuffers buf, DevParams prm, int S, int T) {
  const int b = blockIdx.y;
  const int n = buf.n[b];
  const int warp = threadIdx.x >> 5, lane = lane_id();
  const int chunk = blockIdx.x * kWarpsPerBlock + warp;
  __shared__ unsigned s_run[kWarpsPerBlock][kKeys];
  if (chunk * kChunk >= n) return;
  unsigned* run = s_run[warp];
  const unsigned* row = buf.hist + ((size_t)b * T + chunk) * kKeys;
  for (int t = lane; t < kKeys; t += 32) run[t] = row[t];
  __syncwarp();
  const unsigned lt = (1u << lane) - 1u;
  for (int it = 0; it < kChunk / 32; it++) {
    const int i = chunk * kChunk + it * 32 + lane;
    int ring = -1, sec = -1;
    float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
    if (i < n) {
      const size_t g = (size_t)b * S + i;
      ring = buf.ringid[g];
      sec = buf.sect[g];
      if (ring >= 0 || sec >= 0) p = __ldg(&buf.in[g]);
    }
    unsigned peers = __match_any_sync(0xffffffffu, ring);
    unsigned dst = 0;
    if (ring >= 0) dst = run[ring] + __popc(peers & lt);
    __syncwarp();
    if (ring >= 0 && lane == __ffs(peers) - 1) run[ring] += __popc(peers);
    if (ring >= 0) buf.bpt[(size_t)b * S + dst] = make_float4(p.x, p.y, p.z, __int_as_float(i));
    peers = __match_any_sync(0xffffffffu, sec);
    if (sec >= 0) dst = run[kRingKeys + sec] + __popc(peers & lt);
    __syncwarp();
    if (sec >= 0 && lane == __ffs(peers) - 1) run[kRingKeys + sec] += __popc(peers);
    if (sec >= 0) {
      const float r = star_radius(p.x, p.y);
      buf.spt[(size_t)b * S + dst] = make_float4(r, p.z, __int_as_float(i), 0.f);
    }
    __syncwarp();
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// k_star_sort: one CTA per (sector, scan): sort the sector's points by planar radius (star_shaped_search.cpp:109).
// Tie policy: (r, push_back order) — the reference's introsort order for equal r is unspecified; ties raise F_TIE_SECTOR.
constexpr int kStarSmemKeys = 4096;
__global__ void __launch_bounds__(128) k_star_sort(DevBuffers buf, int S) {
  const int b = blockIdx.y, s = blockIdx.x;
  const ScanTab& tab = buf.tab[b];
  const int base = tab.sect_start[s];
  const int n = tab.sect_start[s + 1] - base;
  if (n <= 0) return;
  const float4* src = buf.spt + (size_t)b * S + base;
  float4* dst = buf.ssorted + (size_t)b * S + base;
  if (n == 1) { if (threadIdx.x == 0) dst[0] = src[0]; return; }
  __shared__ unsigned long long s_keys[kStarSmemKeys];
  const int npad = next_pow2(n);
  unsigned long long* keys = npad <= kStarSmemKeys ? s_keys : buf.sortbuf + 2 * ((size_t)b * S + base);
  for (int t = threadIdx.x; t < npad; t += blockDim.x)
    keys[t] = t < n ? (((unsigned long long)fbits(src[t].x) << 32) | (unsigned)t) : ~0ull;
  __syncthreads();
  cta_bitonic(keys, npad);
  bool tie = false;
  for (int t = threadIdx.x; t < n; t += blockDim.x) {
    const unsigned long long k = keys[t];
    dst[t] = src[(unsigned)k];
    if (t > 0 && (unsigned)(keys[t - 1] >> 32) == (unsigned)(k >> 32)) tie = true;
  }
  if (tie) atomicOr(&buf.out[b].flags, F_TIE_SECTOR);
}

// k_star_scan: one lane per sector walks its radius-sorted points with the reference's running mean / average absolute
// deviation recurrence (star_shaped_search.cpp:112-150) and marks the first edge point.
__global__ void __launch_bounds__(32) k_star_scan(DevBuffers buf, DevParams prm, int S) {
  const int b = blockIdx.y;
  const int s = blockIdx.x * 32 + threadIdx.x;
  if (s >= kSectKeys) return;
  const ScanTab& tab = buf.tab[b];
  const int base = tab.sect_start[s];
  const int n = tab.sect_start[s + 1] - base;
  const float4* pts = buf.ssorted + (size_t)b * S + base;
  const int hit = star_scan_sector(prm, pts, n);
  if (hit >= 0) buf.mark[(size_t)b * S + __float_as_int(pts[hit].z)] = 2;   // star_shaped_search.cpp:146
}

// ---------------------------------------------------------------------------------------------------------------------
// k_ring_detect: one thread per ring-bucket position. Planar range + azimuth (lidar_segmentation.cpp:245-274), the
// x-zero test for which this point is the middle point p2 (x_zero_method.cpp:30-67), the z-zero test centred on it
// (z_zero_method.cpp:21-72), and the curb aggregates blindSpots needs.
__global__ void __launch_bounds__(256) k_ring_detect(DevBuffers buf, DevParams prm, int S) {
  const int b = blockIdx.y;
  const ScanOut& out = buf.out[b];
  const int R = out.n_rings, N = out.n_order;
  __shared__ int s_rs[kRingKeys + 1];
  for (int t = threadIdx.x; t <= kRingKeys; t += blockDim.x) s_rs[t] = out.ring_start[t];
  __syncthreads();
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  const bool act = p < N;
  int k = -1;
  unsigned dbits = 0;
  if (act) {
    k = ring_of(s_rs, R, p);
    const int base = s_rs[k], n = s_rs[k + 1] - base, m = p - base;
    const float4* ring = buf.bpt + (size_t)b * S + base;
    const float4 me = ring[m];
    const float x = me.x, y = me.y;
    const int idx = __float_as_int(me.w);
    float d, az;
    planar_az(x, y, &d, &az);                                           // lidar_segmentation.cpp:245-269
    buf.az[(size_t)b * S + p] = az;
    buf.d2[(size_t)b * S + p] = d;
    dbits = fbits(d);
    int lab = prm.star ? buf.mark[(size_t)b * S + idx] : 0;             // :241-242
    if (prm.x_zero && lab != 2 && xzero_mark(prm, ring, n, m, buf.newY)) lab = 2;   // x_zero_method.cpp:66
    if (prm.z_zero && lab != 2 && zzero_mark(prm, ring, n, m)) lab = 2;             // z_zero_method.cpp:71
    buf.blabel[(size_t)b * S + p] = (unsigned char)lab;
    if (lab == 2 && az >= 0.0f) {            // curb aggregates per (ring, integer-degree bin); NaN azimuths fall out
      const size_t o = ((size_t)b * prm.channels + k) * kDegBins + deg_bin(az);
      atomicMin(&buf.cmin[o], fbits(az));
      atomicMax(&buf.cmax[o], fbits(az));
    }
  }
  // maxDistance[k], lidar_segmentation.cpp:271-274 (warp-aggregated when the whole warp sits in one ring)
  const int k0 = __shfl_sync(0xffffffffu, k, 0);
  if (__all_sync(0xffffffffu, k == k0)) {
    if (k0 >= 0) {
      const unsigned mx = __reduce_max_sync(0xffffffffu, dbits);
      if (lane_id() == 0) atomicMax(&buf.tab[b].maxdist[k0], mx);
    }
  } else if (act) atomicMax(&buf.tab[b].maxdist[k], dbits);
}

// ---------------------------------------------------------------------------------------------------------------------
// blindSpots as tables: see urf_logic.cuh (CurbView, window_reach, covered_by_window).
__global__ void __launch_bounds__(384) k_tables(DevBuffers buf, DevParams prm) {
  const int b = blockIdx.x;
  const ScanOut& out = buf.out[b];
  ScanTab& tab = buf.tab[b];
  const int R = out.n_rings;
  if (R <= 0) return;
  const size_t nb = (size_t)prm.channels * kDegBins;
  CurbView cv{buf.cmin + (size_t)b * nb, buf.cmax + (size_t)b * nb, buf.ne + (size_t)b * prm.channels * (kDegBins + 1)};
  unsigned short* ne = buf.ne + (size_t)b * prm.channels * (kDegBins + 1);
  __shared__ float s_q[4];
  // (1) prefix count of non-empty curb bins per ring
  for (int k = threadIdx.x; k < R; k += blockDim.x) {
    unsigned short run = 0;
    for (int bin = 0; bin < kDegBins; bin++) {
      ne[(size_t)k * (kDegBins + 1) + bin] = run;
      run += cv.cmin[(size_t)k * kDegBins + bin] != 0x7f800000u;
    }
    ne[(size_t)k * (kDegBins + 1) + kDegBins] = run;
  }
  // (2) per-ring arc widths, blind_spots.cpp:65,142
  {
    const float arc = arc_distance(prm, bitsf(tab.maxdist[0]));
    for (int k = threadIdx.x; k < R; k += blockDim.x) tab.A[k] = ring_width(arc, bitsf(tab.maxdist[k]));
  }
  // (3) q1..q4 from the curb points of ring index 1, blind_spots.cpp:13-57
  if (threadIdx.x < 4) {
    const float q = blind_quarter(prm, cv, R, threadIdx.x);
    s_q[threadIdx.x] = q;
    tab.q[threadIdx.x] = q;
  }
  __syncthreads();
  // (4) how many rings each window start accepts: forward (blind_spots.cpp:68-174) and backward (:177-283)
  for (int t = threadIdx.x; t < 2 * kDegBins; t += blockDim.x) {
    const int dir = t / kDegBins, i = t % kDegBins;
    const int reach = window_reach(prm, cv, tab.A, s_q, R, dir, i);
    tab.reach[dir][i] = (unsigned short)reach;
    tab.st[dir][0][i] = (unsigned short)reach;
  }
  __syncthreads();
  // (5) range-max sparse tables over the window starts
  for (int l = 1; l < kStLevels; l++) {
    for (int t = threadIdx.x; t < 2 * kDegBins; t += blockDim.x) {
      const int dir = t / kDegBins, i = t % kDegBins;
      const int j = i + (1 << (l - 1));
      const unsigned short a = tab.st[dir][l - 1][i];
      const unsigned short c = j < kDegBins ? tab.st[dir][l - 1][j] : (unsigned short)0;
      tab.st[dir][l][i] = a > c ? a : c;
    }
    __syncthreads();
  }
}

// k_label: final label per ring-bucket position, scattered back to input order; counts; first non-road ring per bin.
__global__ void __launch_bounds__(256) k_label(DevBuffers buf, DevParams prm, int S) {
  const int b = blockIdx.y;
  ScanOut& out = buf.out[b];
  ScanTab& tab = buf.tab[b];
  const int R = out.n_rings, N = out.n_order;
  __shared__ int s_rs[kRingKeys + 1];
  __shared__ int s_road, s_curb;
  for (int t = threadIdx.x; t <= kRingKeys; t += blockDim.x) s_rs[t] = out.ring_start[t];
  if (threadIdx.x == 0) { s_road = 0; s_curb = 0; }
  __syncthreads();
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  int lab = -1;
  if (p < N) {
    const size_t g = (size_t)b * S + p;
    const int k = ring_of(s_rs, R, p);
    const float a = buf.az[g];
    lab = buf.blabel[g];
    if (lab != 2 && covered_by_window(prm, tab, k, a)) lab = 1;
    buf.blabel[g] = (unsigned char)lab;
    const int idx = __float_as_int(buf.bpt[g].w);
    buf.label[(size_t)b * S + idx] = lab;
    if (lab != 1 && a >= 0.0f) {                   // lidar_segmentation.cpp:318: non-road point in bin [i, i+1)
      const int bin = deg_bin(a);
      if (tab.cut[bin] > k) atomicMin(&tab.cut[bin], k);
    }
  }
  const unsigned br = __ballot_sync(0xffffffffu, lab == 1), bc = __ballot_sync(0xffffffffu, lab == 2);
  if (lane_id() == 0) { if (br) atomicAdd(&s_road, __popc(br)); if (bc) atomicAdd(&s_curb, __popc(bc)); }
  __syncthreads();
  if (threadIdx.x == 0) { if (s_road) atomicAdd(&out.n_road, s_road); if (s_curb) atomicAdd(&out.n_curb, s_curb); }
}

// Marker candidate vertices, lidar_segmentation.cpp:305-351, as three order-independent passes over the ring buckets:
//   k_cutkey: first non-road point (azimuth, bucket position) of the cut ring of every bin
//   k_dmax  : farthest candidate road point per bin (candidates: road points scanned before the first non-road point)
//   k_best  : first candidate (ring, azimuth, position order) that reaches that distance (`d > maxDistanceRoad` is strict)
__device__ __forceinline__ bool marker_point(const DevBuffers& buf, const ScanOut& out, const int* s_rs, int b, int S,
                                             int p, int& k, int& lab, int& bin, unsigned& abits) {
  if (p >= out.n_order) return false;
  const size_t g = (size_t)b * S + p;
  const float a = buf.az[g];
  if (!(a >= 0.0f)) return false;
  abits = fbits(a);
  bin = deg_bin(a);
  k = ring_of(s_rs, out.n_rings, p);
  lab = buf.blabel[g];
  return true;
}

__global__ void __launch_bounds__(256) k_cutkey(DevBuffers buf, int S) {
  const int b = blockIdx.y;
  const ScanOut& out = buf.out[b];
  ScanTab& tab = buf.tab[b];
  __shared__ int s_rs[kRingKeys + 1];
  for (int t = threadIdx.x; t <= kRingKeys; t += blockDim.x) s_rs[t] = out.ring_start[t];
  __syncthreads();
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  int k, lab, bin; unsigned ab;
  if (!marker_point(buf, out, s_rs, b, S, p, k, lab, bin, ab)) return;
  if (lab != 1 && k == tab.cut[bin]) atomicMin(&tab.cutkey[bin], cut_key(ab, p));
}

__global__ void __launch_bounds__(256) k_dmax(DevBuffers buf, int S) {
  const int b = blockIdx.y;
  const ScanOut& out = buf.out[b];
  ScanTab& tab = buf.tab[b];
  __shared__ int s_rs[kRingKeys + 1];
  for (int t = threadIdx.x; t <= kRingKeys; t += blockDim.x) s_rs[t] = out.ring_start[t];
  __syncthreads();
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  int k, lab, bin; unsigned ab;
  if (!marker_point(buf, out, s_rs, b, S, p, k, lab, bin, ab)) return;
  if (marker_candidate(tab, k, lab, bin, ab, p)) {
    const unsigned d = fbits(buf.d2[(size_t)b * S + p]);                 // :327 (same value as the ring's planar range)
    if (tab.dmax[bin] < d) atomicMax(&tab.dmax[bin], d);
  }
}

__global__ void __launch_bounds__(256) k_best(DevBuffers buf, int S) {
  const int b = blockIdx.y;
  const ScanOut& out = buf.out[b];
  ScanTab& tab = buf.tab[b];
  __shared__ int s_rs[kRingKeys + 1];
  for (int t = threadIdx.x; t <= kRingKeys; t += blockDim.x) s_rs[t] = out.ring_start[t];
  __syncthreads();
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  int k, lab, bin; unsigned ab;
  if (!marker_point(buf, out, s_rs, b, S, p, k, lab, bin, ab)) return;
  if (marker_candidate(tab, k, lab, bin, ab, p)) {
    const unsigned d = fbits(buf.d2[(size_t)b * S + p]);
    if (d != 0u && d == tab.dmax[bin])                                   // :329 `d > maxDistanceRoad`, initial 0
      atomicMin(&tab.best[bin], best_key(k, ab, p));
  }
}

// k_verts: compact the per-bin winners in bin order into markerPointsArray (lidar_segmentation.cpp:343-350).
__global__ void __launch_bounds__(384) k_verts(DevBuffers buf, int S) {
  const int b = blockIdx.x;
  ScanOut& out = buf.out[b];
  const ScanTab& tab = buf.tab[b];
  __shared__ int s_wsum[12];
  const int i = threadIdx.x;
  const bool has = i < kDegBins && tab.best[i] != ~0ull;
  const unsigned bal = __ballot_sync(0xffffffffu, has);
  const int warp = i >> 5, lane = lane_id();
  if (lane == 0) s_wsum[warp] = __popc(bal);
  __syncthreads();
  int off = 0, total = 0;
  for (int w = 0; w < 12; w++) { if (w < warp) off += s_wsum[w]; total += s_wsum[w]; }
  if (has) {
    const int slot = off + __popc(bal & ((1u << lane) - 1u));
    const int p = (int)(tab.best[i] & 0xffffffull);
    const float4 q = buf.bpt[(size_t)b * S + p];
    out.vert[slot][0] = q.x; out.vert[slot][1] = q.y; out.vert[slot][2] = q.z;
    out.vert[slot][3] = tab.cut[i] != 0x7fffffff ? 1.0f : 0.0f;         // redPoints, :320,348
  }
  if (i == 0) out.n_vert = total;
}

// ---------------------------------------------------------------------------------------------------------------------
// k_sort_rings (only when the emission order is requested): per-ring sort by azimuth, lidar_segmentation.cpp:289-291.
// Tie policy: (azimuth, input order) — the reference's Lomuto quicksort is unstable; ties raise F_TIE_AZIMUTH.
constexpr int kRingSmemKeys = 8192;
__global__ void __launch_bounds__(256) k_sort_rings(DevBuffers buf, int S) {
  extern __shared__ unsigned long long s_rkeys[];
  const int b = blockIdx.y, k = blockIdx.x;
  ScanOut& out = buf.out[b];
  if (k >= out.n_rings) return;
  const int base = out.ring_start[k], n = out.ring_start[k + 1] - base;
  if (n <= 0) return;
  const size_t g0 = (size_t)b * S + base;
  const int npad = next_pow2(n < 2 ? 2 : n);
  unsigned long long* keys = npad <= kRingSmemKeys ? s_rkeys : buf.sortbuf + 2 * g0;
  for (int t = threadIdx.x; t < npad; t += blockDim.x)
    keys[t] = t < n ? (((unsigned long long)fbits(buf.az[g0 + t]) << 32) | (unsigned)t) : ~0ull;
  __syncthreads();
  cta_bitonic(keys, npad);
  bool tie = false;
  for (int t = threadIdx.x; t < n; t += blockDim.x) {
    const unsigned long long key = keys[t];
    buf.order[g0 + t] = __float_as_int(buf.bpt[g0 + (unsigned)key].w);
    if (t > 0 && (unsigned)(keys[t - 1] >> 32) == (unsigned)(key >> 32)) tie = true;
  }
  if (tie) atomicOr(&out.flags, F_TIE_AZIMUTH);
}

// Device-side evaluation of the emulated libm (test hook: urf_test_math).
__global__ void k_test_math(const float* a, const float* bb, float* o, int n, int which) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float r;
  switch (which) {
    case 0: r = urfm::asinf_glibc(a[i]); break;
    case 1: r = urfm::acosf_glibc(a[i]); break;
    case 2: r = urfm::atan2f_glibc(a[i], bb[i]); break;
    default: r = urfm::atanf_glibc(a[i]); break;
  }
  o[i] = r;
}

}  // namespace urf
