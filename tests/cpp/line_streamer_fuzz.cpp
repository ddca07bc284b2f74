// Fuzz of va::LineStreamer (versalignlib_b200/csrc/va_line_streamer.h) against memcpy: random piece lengths (short pieces,
// pieces around half the line buffer and around its size, pieces of 32 000 bytes), random destination alignment, guard
// bytes on both sides of the destination.  Prints "ok <cases>" or the first failure.  Built and run by
// tests/test_line_streamer.py with the host compiler; no GPU.
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>

#include "va_line_streamer.h"

int main(int argc, char **argv) {
    const int cases = argc > 1 ? atoi(argv[1]) : 2000;
    std::mt19937_64 rng(12345);
    const size_t kinds[] = {1, 3, 17, 63, 64, 65, 100, 150, 151, 250, 1000, 4000, 4095, 4096, 4097, 5000, 8128, 8129, 8191, 8192, 8193, 12000, 32000};
    for (int c = 0; c < cases; ++c) {
        const size_t len = (rng() % 4 == 0) ? 1 + rng() % 9000 : kinds[rng() % (sizeof(kinds) / sizeof(kinds[0]))];
        const int pieces = 1 + (int)(rng() % (len > 4096 ? 12 : 300));
        const size_t align = rng() % 64, guard = 128;
        std::vector<char> src((size_t)pieces * len), want((size_t)pieces * len + align + 2 * guard + 64, (char)0x5A);
        for (auto &b : src) b = (char)rng();
        std::vector<char> got(want);
        char *base = got.data();
        base += (64 - (reinterpret_cast<uintptr_t>(base) & 63)) & 63;  // line-aligned, then the test's own misalignment
        const size_t lead = (size_t)(base - got.data()) + guard + align;
        memcpy(want.data() + lead, src.data(), src.size());
        {
            va::LineStreamer w(got.data() + lead);
            for (int p = 0; p < pieces; ++p) w.append(src.data() + (size_t)p * len, len);
            w.finish();
        }
        if (memcmp(want.data(), got.data(), want.size()) != 0) {
            size_t at = 0;
            while (want[at] == got[at]) ++at;
            printf("FAIL case %d: len %zu pieces %d align %zu first difference at byte %zd of the destination\n", c, len, pieces, align,
                   (ssize_t)at - (ssize_t)lead);
            return 1;
        }
    }
    printf("ok %d\n", cases);
    return 0;
}
