// Host copy (staging buffer -> destination array) with non-temporal stores of 16 / 32 / 64 bytes, N threads.
// build: g++ -O3 -std=c++17 -pthread -o hostcopy_bench hostcopy_bench.cpp ; run: ./hostcopy_bench [MB] [threads...]
#include <immintrin.h>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>
static void copy_sse2(char* d, const char* s, size_t n) {
  for (size_t i = 0; i < n; i += 16) _mm_stream_si128((__m128i*)(d + i), _mm_loadu_si128((const __m128i*)(s + i)));
  _mm_sfence();
}
__attribute__((target("avx2"))) static void copy_avx2(char* d, const char* s, size_t n) {
  for (size_t i = 0; i < n; i += 32) _mm256_stream_si256((__m256i*)(d + i), _mm256_loadu_si256((const __m256i*)(s + i)));
  _mm_sfence();
}
__attribute__((target("avx512f"))) static void copy_avx512(char* d, const char* s, size_t n) {
  for (size_t i = 0; i < n; i += 64) _mm512_stream_si512((__m512i*)(d + i), _mm512_loadu_si512((const void*)(s + i)));
  _mm_sfence();
}
static void copy_memcpy(char* d, const char* s, size_t n) { memcpy(d, s, n); }
int main(int argc, char** argv) {
  const size_t mb = argc > 1 ? atoll(argv[1]) : 1024, n = mb << 20;
  char* src = (char*)aligned_alloc(4096, n);
  char* dst = (char*)aligned_alloc(4096, n);
  memset(src, 1, n);
  memset(dst, 0, n);
  struct V { const char* name; void (*fn)(char*, const char*, size_t); bool ok; };
  V vs[] = {{"sse2 16 B", copy_sse2, true}, {"avx2 32 B", copy_avx2, (bool)__builtin_cpu_supports("avx2")},
            {"avx512 64 B", copy_avx512, (bool)__builtin_cpu_supports("avx512f")}, {"memcpy", copy_memcpy, true}};
  std::vector<int> ths;
  for (int i = 2; i < argc; ++i) ths.push_back(atoi(argv[i]));
  if (ths.empty()) ths = {1, 4, 8, 16};
  for (auto& v : vs) {
    if (!v.ok) { printf("%-12s not supported by this CPU\n", v.name); continue; }
    for (int nt : ths) {
      double best = 1e9;
      for (int rep = 0; rep < 3; ++rep) {
        auto t0 = std::chrono::steady_clock::now();
        std::vector<std::thread> pool;
        for (int t = 0; t < nt; ++t) {
          const size_t lo = (n / 64 * t / nt) * 64, hi = (n / 64 * (t + 1) / nt) * 64;
          pool.emplace_back(v.fn, dst + lo, src + lo, hi - lo);
        }
        for (auto& th : pool) th.join();
        const double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        best = s < best ? s : best;
      }
      printf("%-12s %2d threads: %6.1f GB/s\n", v.name, nt, n / best / 1e9);
    }
  }
  return 0;
}
