// tests/kat/stdsort_check.cpp — urban_road_filter_b200/csrc/urf_stdsort.cuh (the restatement of libstdc++'s std::sort that
// the star search's tie path runs on the device) against the REAL std::sort of this toolchain, on the element type and
// comparator the reference uses (star_shaped_search.cpp:22-25,109: `polar` records compared by r only).
// Arrays: random radii drawn from few distinct values (many ties), sorted / reversed / organ-pipe / constant inputs, and
// "median-of-three killer" sequences that drive introsort into its heapsort fallback. usage: stdsort_check [rounds]
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <vector>
static long g_heap = 0;
#define URF_STDSORT_COUNT_HEAP g_heap
#include "../../urban_road_filter_b200/csrc/urf_stdsort.cuh"

struct polar { int id; float r; float fi; };        // data_structures.hpp: the reference's record (id, r, fi)
static bool ptcmpr(polar a, polar b) { return (a.r < b.r); }

static bool check(const std::vector<float>& r) {
  const long n = (long)r.size();
  std::vector<polar> ref(n);
  std::vector<urfsort::El> mine(n);
  for (long i = 0; i < n; i++) {
    ref[i] = polar{(int)i, r[i], 0.f};
    uint32_t b; memcpy(&b, &r[i], 4);
    mine[i] = ((urfsort::El)b << 32) | (uint32_t)i;
  }
  std::sort(ref.begin(), ref.end(), ptcmpr);
  urfsort::std_sort(mine.data(), n);
  for (long i = 0; i < n; i++) if (ref[i].id != (int)(uint32_t)mine[i]) { fprintf(stderr, "n=%ld: position %ld holds id %d, std::sort put %d there\n", n, i, (int)(uint32_t)mine[i], ref[i].id); return false; }
  return true;
}

// Musser's median-of-3 killer permutation: forces quadratic partitioning, i.e. the depth limit and the heapsort branch
static std::vector<float> killer(int n) {
  std::vector<float> v(n);
  const int k = n / 2;
  for (int i = 0; i < k; i++) { if (i % 2 == 0) v[i] = (float)(i + 1); else v[i] = (float)(k + i + (k % 2 == 0 ? 0 : 1)); v[k + i] = (float)(2 * (i + 1)); }
  if (n % 2) v[n - 1] = (float)n;
  return v;
}

int main(int argc, char** argv) {
  const int rounds = argc > 1 ? atoi(argv[1]) : 3000;
  std::mt19937 g(12345);
  long cases = 0;
  for (int round = 0; round < rounds; round++) {
    const int n = round < 40 ? round : (int)(g() % (round % 50 == 0 ? 20000 : 1500));
    const int distinct = 1 + (int)(g() % (round % 3 == 0 ? 4 : (round % 3 == 1 ? 40 : 4000)));
    std::vector<float> r(n);
    for (float& x : r) x = 1.0f + 0.25f * (float)(g() % distinct);
    switch (round % 7) {
      case 1: std::sort(r.begin(), r.end()); break;
      case 2: std::sort(r.begin(), r.end()); std::reverse(r.begin(), r.end()); break;
      case 3: { std::sort(r.begin(), r.end()); std::reverse(r.begin() + n / 2, r.end()); } break;      // organ pipe
      case 4: for (int i = 0; i + 1 < n; i += 2) std::swap(r[i], r[i + 1]); break;
      default: break;
    }
    if (!check(r)) return 1;
    cases++;
  }
  for (int n : {17, 33, 64, 100, 257, 1000, 4096, 10000, 65536}) {
    std::vector<float> k = killer(n);
    if (!check(k)) return 1;
    for (float& x : k) x = (float)((int)x / 3);        // the same shape with ties
    if (!check(k)) return 1;
    std::vector<float> c(n, 2.5f);                     // all equal
    if (!check(c)) return 1;
    cases += 3;
  }
  printf("stdsort checked=%ld mismatches=0 heapsort_fallbacks=%ld\n", cases, g_heap);
  return 0;
}
