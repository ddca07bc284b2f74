// va_device.cuh -- small device-side helpers shared by the kernels.
#ifndef VA_DEVICE_CUH
#define VA_DEVICE_CUH

#include <cuda_runtime.h>
#include <stdint.h>

namespace va {

// PTX prmt.b32 in its default mode: every selector nibble picks one of the 8 source bytes
// {b (bytes 4..7), a (bytes 0..3)}; bit 3 of the nibble replicates that byte's sign bit
// instead -- which is how an 8-bit table entry becomes a sign-extended 16-bit lane in one
// instruction.  (__byte_perm() masks the selector with 0x7777 and loses that mode.)
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {
    uint32_t r;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(sel));
    return r;
}

// Start of raw sequence `pair`: fixed stride, or through the caller's offsets (packed entry points).
__device__ __forceinline__ const uint8_t *seq_ptr(const uint8_t *raw, const int64_t *off, int pair, int stride) {
    return off ? raw + (off[pair] - off[0]) : raw + (size_t)pair * (size_t)stride;
}
// One past the last byte of the whole raw buffer of an n-pair chunk.
__device__ __forceinline__ const uint8_t *seq_end(const uint8_t *raw, const int64_t *off, int n, int stride) {
    return off ? raw + (off[n] - off[0]) : raw + (size_t)n * (size_t)stride;
}

}  // namespace va
#endif
