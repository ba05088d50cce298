// tests/kat/math_sweep.cpp — known-answer sweep of urf_math.cuh (host build) against the container's glibc libm.
// usage: math_sweep <stride> <threads> <n_random_atan2>
//   asinf/acosf: every `stride`-th float bit pattern of [-1, 1] plus out-of-range/NaN probes
//   atanf      : every `stride`-th of all 2^32 bit patterns
//   atan2f     : n_random pairs (several distributions) + a grid of special values
// Prints one line per function: "<name> checked=<n> mismatches=<m>"; exit code 1 if any mismatch.
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>
#include <atomic>
#include <random>
#include "../../urban_road_filter_b200/csrc/urf_math.cuh"
#include "../../urban_road_filter_b200/csrc/urf_logic.cuh"
#include "../../urban_road_filter_b200/csrc/urf_host.hpp"

static inline uint32_t bits(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline float fl(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
static inline bool same(float a, float b) { return bits(a) == bits(b) || (std::isnan(a) && std::isnan(b)); }

template <class F, class G>
static uint64_t sweep1(const char* name, uint64_t lo, uint64_t hi, uint64_t stride, int threads, F mine, G ref, uint64_t* checked) {
  std::atomic<uint64_t> bad{0}, cnt{0};
  std::vector<std::thread> th;
  for (int t = 0; t < threads; t++) {
    th.emplace_back([&, t]() {
      uint64_t b = 0, c = 0;
      for (uint64_t u = lo + (uint64_t)t * stride; u < hi; u += stride * threads) {
        float x = fl((uint32_t)u);
        float a = mine(x), r = ref(x);
        if (!same(a, r)) { if (b < 3) fprintf(stderr, "%s(%a [%08x]) = %a, libm %a\n", name, x, (uint32_t)u, a, r); b++; }
        c++;
      }
      bad += b; cnt += c;
    });
  }
  for (auto& x : th) x.join();
  *checked += cnt;
  return bad;
}

int main(int argc, char** argv) {
  uint64_t stride = argc > 1 ? strtoull(argv[1], 0, 10) : 1;
  int threads = argc > 2 ? atoi(argv[2]) : 8;
  uint64_t nrand = argc > 3 ? strtoull(argv[3], 0, 10) : 200000000ull;
  int rc = 0;
  {
    // [-1,1] = bit patterns [0, 0x3f800000] and [0x80000000, 0xbf800000]; probe a little beyond too
    uint64_t c = 0, b = 0;
    b += sweep1("asinf", 0, 0x3f800000ull + 64, stride, threads, urfm::asinf_glibc, (float (*)(float))asinf, &c);
    b += sweep1("asinf", 0x80000000ull, 0xbf800000ull + 64, stride, threads, urfm::asinf_glibc, (float (*)(float))asinf, &c);
    b += sweep1("asinf", 0x7f7fff00ull, 0x7fc00010ull, 1, 1, urfm::asinf_glibc, (float (*)(float))asinf, &c);
    printf("asinf checked=%llu mismatches=%llu\n", (unsigned long long)c, (unsigned long long)b);
    rc |= b != 0;
    c = b = 0;
    b += sweep1("acosf", 0, 0x3f800000ull + 64, stride, threads, urfm::acosf_glibc, (float (*)(float))acosf, &c);
    b += sweep1("acosf", 0x80000000ull, 0xbf800000ull + 64, stride, threads, urfm::acosf_glibc, (float (*)(float))acosf, &c);
    b += sweep1("acosf", 0x7f7fff00ull, 0x7fc00010ull, 1, 1, urfm::acosf_glibc, (float (*)(float))acosf, &c);
    printf("acosf checked=%llu mismatches=%llu\n", (unsigned long long)c, (unsigned long long)b);
    rc |= b != 0;
    c = b = 0;
    b += sweep1("atanf", 0, 0x100000000ull, stride, threads, urfm::atanf_glibc, (float (*)(float))atanf, &c);
    printf("atanf checked=%llu mismatches=%llu\n", (unsigned long long)c, (unsigned long long)b);
    rc |= b != 0;
    c = b = 0;
    {   // div_pi((double)f) against the IEEE double quotient, all 64 bits, every `stride`-th finite float
      std::atomic<uint64_t> bad{0}, cnt{0};
      std::vector<std::thread> th;
      for (int t = 0; t < threads; t++) th.emplace_back([&, t]() {
        uint64_t bb = 0, cc = 0;
        for (uint64_t u = (uint64_t)t * stride; u < 0x100000000ull; u += stride * threads) {
          const float f = fl((uint32_t)u);
          if (!std::isfinite(f)) continue;
          const double mine = urfm::div_pi((double)f), ref = (double)f / URF_PI_D;
          if (memcmp(&mine, &ref, 8)) { if (bb < 3) fprintf(stderr, "div_pi(%a) = %a, IEEE %a\n", (double)f, mine, ref); bb++; }
          if (!(u & 0x80000000ull)) {                       // the variant without the sign-of-zero test, on +0 and above
            const double nn = urfm::div_pi_nonneg((double)f);
            if (memcmp(&nn, &ref, 8)) { if (bb < 3) fprintf(stderr, "div_pi_nonneg(%a) = %a, IEEE %a\n", (double)f, nn, ref); bb++; }
          }
          cc++;
        }
        bad += bb; cnt += cc;
      });
      for (auto& x : th) x.join();
      b += bad; c += cnt;
    }
    printf("div_pi checked=%llu mismatches=%llu\n", (unsigned long long)c, (unsigned long long)b);
    rc |= b != 0;
  }
  {
    std::atomic<uint64_t> bad{0}, cnt{0};
    std::vector<std::thread> th;
    for (int t = 0; t < threads; t++) {
      th.emplace_back([&, t]() {
        std::mt19937_64 g(1234 + t);
        std::uniform_real_distribution<float> U(-100.f, 100.f), S(-1e-3f, 1e-3f);
        uint64_t b = 0, c = 0;
        for (uint64_t i = t; i < nrand; i += threads) {
          float y, x;
          switch (i & 7) {
            case 0: y = U(g); x = U(g); break;
            case 1: y = S(g); x = U(g); break;                          // tiny |y/x|: the f==360 corner of star search
            case 2: y = U(g); x = S(g); break;
            case 3: y = fl((uint32_t)g()); x = fl((uint32_t)g()); break;  // arbitrary bit patterns
            case 4: y = U(g) * 1e-6f; x = U(g); break;
            case 5: { float r = U(g); y = r; x = r * (1.0f + S(g)); } break;
            case 6: y = U(g); x = 1.0f; break;
            default: y = -fabsf(S(g)) * 1e-4f; x = fabsf(U(g)); break;
          }
          float a = urfm::atan2f_glibc(y, x), r = atan2f(y, x);
          if (!same(a, r)) { if (b < 3) fprintf(stderr, "atan2f(%a, %a) = %a, libm %a\n", y, x, a, r); b++; }
          c++;
        }
        bad += b; cnt += c;
      });
    }
    for (auto& x : th) x.join();
    const float sp[] = {0.f, -0.f, 1.f, -1.f, INFINITY, -INFINITY, NAN, 1e-45f, -1e-45f, 3.4e38f, -3.4e38f, 1e-30f, 2.f, 0.5f};
    uint64_t b = 0, c = 0;
    for (float y : sp) for (float x : sp) {
      float a = urfm::atan2f_glibc(y, x), r = atan2f(y, x);
      if (!same(a, r)) { fprintf(stderr, "atan2f(%a, %a) = %a, libm %a\n", y, x, a, r); b++; }
      c++;
    }
    bad += b; cnt += c;
    printf("atan2f checked=%llu mismatches=%llu\n", (unsigned long long)cnt.load(), (unsigned long long)bad.load());
    rc |= bad != 0;
  }
  {
    // star_sector_fast (urf_logic.cuh): every decided point must get the sector the reference's expression gives with the
    // REAL libm (star_shaped_search.cpp:166-171); undecided points (-1) take the exact path in the product. Random points
    // in several distributions plus points placed within +-2e-3 degrees of every sector boundary at many radii.
    float bd[urf::kSectKeys], bo[urf::kSectKeys], Kfi; unsigned char byx[urf::kSectKeys];
    urf::host_beam_init(bd, bo, byx, &Kfi);
    auto exact = [&](float x, float y) {
      float fi = atan2f(y, x);
      if (fi < 0) fi = (float)((double)fi + 2 * M_PI);
      int f = (int)(fi * Kfi);
      if (f >= urf::kSectKeys || f < 0) f = 0;
      return f;
    };
    std::atomic<uint64_t> bad{0}, cnt{0}, undecided{0};
    std::vector<std::thread> th;
    for (int t = 0; t < threads; t++) {
      th.emplace_back([&, t]() {
        std::mt19937_64 g(99 + t);
        std::uniform_real_distribution<float> U(-100.f, 100.f), S(-1e-3f, 1e-3f);
        std::uniform_real_distribution<double> R(0.3, 250.0), E(-2e-3, 2e-3);
        uint64_t b = 0, c = 0, u = 0;
        for (uint64_t i = t; i < nrand; i += threads) {
          float y, x;
          switch (i & 7) {
            case 0: case 1: y = U(g); x = U(g); break;
            case 2: y = S(g); x = U(g); break;
            case 3: y = fl((uint32_t)g()); x = fl((uint32_t)g()); break;      // arbitrary bit patterns (NaN, inf, denormals)
            case 4: y = U(g); x = S(g) * 1e-3f; break;
            default: {                                                        // next to a sector boundary
              const double deg = (double)(g() % 361) + E(g), r = R(g);
              x = (float)(r * cos(deg * M_PI / 180)); y = (float)(r * sin(deg * M_PI / 180));
            } break;
          }
          const int fast = urf::star_sector_fast(x, y);
          if (fast < 0) { u++; c++; continue; }
          if (!std::isfinite(x) || !std::isfinite(y)) { if (b < 3) fprintf(stderr, "sector_fast decided a non-finite point (%a, %a)\n", x, y); b++; }
          else if (fast != exact(x, y)) { if (b < 3) fprintf(stderr, "sector_fast(%a, %a) = %d, reference %d\n", x, y, fast, exact(x, y)); b++; }
          c++;
        }
        bad += b; cnt += c; undecided += u;
      });
    }
    for (auto& x : th) x.join();
    printf("sector checked=%llu mismatches=%llu undecided=%llu\n", (unsigned long long)cnt.load(), (unsigned long long)bad.load(),
           (unsigned long long)undecided.load());
    rc |= bad != 0;
  }
  return rc;
}
