#include <chrono>
#include <cstdio>
#include <cstring>
#include <thread>
#include <vector>
#include <cstdlib>
using Clock = std::chrono::steady_clock;
#include <malloc.h>
int main(int argc, char** argv) {
  if (argc > 2) { mallopt(M_TRIM_THRESHOLD, 1<<30); mallopt(M_TOP_PAD, 64<<20); mallopt(M_MMAP_THRESHOLD, 1<<30); }
  int n = 1000000, L = 300, T = argc > 1 ? atoi(argv[1]) : 8;
  std::vector<char> src((size_t)n * L * 2, 'A');
  std::vector<char*> a(n), b(n);
  for (int rep = 0; rep < 3; ++rep) {
    for (int mode = 0; mode < 3; ++mode) {
      auto t0 = Clock::now();
      std::vector<std::thread> th;
      for (int t = 0; t < T; ++t) th.emplace_back([&, t] {
        int lo = (long long)n * t / T, hi = (long long)n * (t + 1) / T;
        for (int i = lo; i < hi; ++i) {
          if (mode == 0) { a[i] = new char[L]; b[i] = new char[L]; }
          else if (mode == 1) { a[i] = new char[L]; b[i] = new char[L]; memcpy(a[i] + 145, &src[(size_t)i * L + 145], 155); memcpy(b[i] + 145, &src[(size_t)(n + i) * L + 145], 155); }
          else { memcpy(a[i] + 145, &src[(size_t)i * 320], 155); memcpy(b[i] + 145, &src[(size_t)i * 320 + 160], 155); }
        }
      });
      for (auto& x : th) x.join();
      double ms = std::chrono::duration<double, std::milli>(Clock::now() - t0).count();
      printf("rep %d mode %d (%s): %.1f ms\n", rep, mode, mode == 0 ? "alloc only" : mode == 1 ? "alloc+copy strided" : "copy only compact", ms);
      if (mode != 1) { if (mode == 0) for (int i = 0; i < n; ++i) { delete[] a[i]; delete[] b[i]; } }
    }
    for (int i = 0; i < n; ++i) { delete[] a[i]; delete[] b[i]; }
  }
}
