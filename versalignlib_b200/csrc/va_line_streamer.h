// va_line_streamer.h -- host-side helper of the staging pipeline (va_cabi.cu); plain C++ so that a CPU test can fuzz it
// (tests/test_line_streamer.py).
#ifndef VA_LINE_STREAMER_H
#define VA_LINE_STREAMER_H

#include <cstddef>
#include <cstdint>
#include <cstring>
#if defined(__SSE2__)
#include <emmintrin.h>
#endif

namespace va {

// Sequential writer into the pinned staging block that never reads what it overwrites: the bytes are collected in a small
// cache-resident buffer and leave as whole 64-byte lines with non-temporal stores.  A plain memcpy of 150-byte pieces
// makes the core fetch every destination line first (read for ownership); the legacy boundary is bound by the host's
// memory system, and that fetch was a third of the gather's traffic.
#if defined(__SSE2__)
class LineStreamer {
  public:
    explicit LineStreamer(char *dst) : dst_(dst) {}
    void append(const char *src, size_t len) {
        if (len > sizeof(buf_) / 2) {  // long sequences: what is pending goes out first, then a plain copy
            drain();
            memcpy(dst_, src, len);
            dst_ += len;
            return;
        }
        if (fill_ + len > sizeof(buf_)) flush();
        memcpy(buf_ + fill_, src, len);
        fill_ += len;
    }
    void finish() {
        drain();
        _mm_sfence();  // the copy engine reads these lines next
    }

  private:
    void drain() {  // everything collected so far reaches the destination (whole lines streamed, the rest copied)
        flush();
        memcpy(dst_, buf_, fill_);
        dst_ += fill_;
        fill_ = 0;
    }
    void flush() {
        size_t pos = 0;
        const size_t head = (size_t)(-(intptr_t)reinterpret_cast<uintptr_t>(dst_)) & 63u;
        if (head && fill_ >= head) {  // up to the first line boundary of the destination
            memcpy(dst_, buf_, head);
            dst_ += head;
            pos = head;
        } else if (head) {
            return;
        }
        for (; pos + 64 <= fill_; pos += 64, dst_ += 64) {
            const __m128i a = _mm_loadu_si128(reinterpret_cast<const __m128i *>(buf_ + pos));
            const __m128i b = _mm_loadu_si128(reinterpret_cast<const __m128i *>(buf_ + pos + 16));
            const __m128i c = _mm_loadu_si128(reinterpret_cast<const __m128i *>(buf_ + pos + 32));
            const __m128i d = _mm_loadu_si128(reinterpret_cast<const __m128i *>(buf_ + pos + 48));
            _mm_stream_si128(reinterpret_cast<__m128i *>(dst_), a);
            _mm_stream_si128(reinterpret_cast<__m128i *>(dst_ + 16), b);
            _mm_stream_si128(reinterpret_cast<__m128i *>(dst_ + 32), c);
            _mm_stream_si128(reinterpret_cast<__m128i *>(dst_ + 48), d);
        }
        memmove(buf_, buf_ + pos, fill_ - pos);
        fill_ -= pos;
    }
    char *dst_;
    size_t fill_ = 0;
    alignas(64) char buf_[8192];
};
#else
class LineStreamer {
  public:
    explicit LineStreamer(char *dst) : dst_(dst) {}
    void append(const char *src, size_t len) {
        memcpy(dst_, src, len);
        dst_ += len;
    }
    void finish() {}

  private:
    char *dst_;
};
#endif

}  // namespace va
#endif
