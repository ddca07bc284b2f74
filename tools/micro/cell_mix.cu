// cell_mix.cu -- which NW-align cell formulation sweeps a 30-column register strip fastest on sm_100a?
// Registers only (row tables from shared memory, direction words to a per-thread global slot like the real kernel), so
// the numbers are the instruction-mix ceiling of each form, 4 x 128-thread blocks per SM like fill_nw_kernel<ALIGN>.
//   KIND 0  planes by predicates (what va_nw.cu ships): PRMT, VIMNMX.pred, add, VIMNMX.pred, 4 predicated FADD
//   KIND 1  in-band tags: values are 4V + tag, PRMT, 2 VIADDMNMX, LOP3 (clean), tags banked by 2 IMAD (x4 Horner)
//   KIND 2  in-band tags, banked on the integer pipe (LOP3 + SHF/LOP3)  -- control
//   KIND 3  in-band tags, nothing banked -- ceiling of the 4-instruction integer form
//   KIND 5, 6  KIND 1 with the multiply-adds as inline PTX (ptxas otherwise splits some into SHL + IADD3)
//   KIND 4  score form (PRMT, VIMNMX, VIADDMNMX) -- ceiling of the 3-instruction form
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o cell_mix cell_mix.cu ; run: ./cell_mix
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t s) { uint32_t r; asm("prmt.b32 %0,%1,%2,%3;" : "=r"(r) : "r"(a), "r"(b), "r"(s)); return r; }
constexpr uint32_t NEG2 = 0x80008000u;

template <int KIND>
__global__ void __maxnreg__(128) k(int rows, uint32_t seed, float one, const uint32_t *selsrc, uint4 *out, unsigned long long *cyc) {
    constexpr int TW = 30;
    __shared__ uint2 s_T2[64];
    if (threadIdx.x < 64) s_T2[threadIdx.x] = make_uint2(0x05050508u + (threadIdx.x & 3) * 0x01000000u + seed, 0x05080505u + seed);
    __syncthreads();
    uint32_t sel[TW], H[TW];
#pragma unroll
    for (int q = 0; q < TW; ++q) {
        const uint32_t v = selsrc[q * 128 + threadIdx.x], fa = v & 3, fb = (v >> 2) & 3;  // opaque to the compiler
        sel[q] = fa | ((fa | 8u) << 4) | ((fb | 4u) << 8) | ((fb | 12u) << 12);
        H[q] = 0;
    }
    uint32_t diag_next = 0;
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, nthreads = (size_t)gridDim.x * blockDim.x;
    uint4 *dp = out + tid;
    const unsigned long long t0 = clock64();
    #pragma unroll 2
    for (int r = 0; r < rows; ++r) {
        const uint2 tt = s_T2[(r * 7 + threadIdx.x) & 63];
        uint32_t left = (uint32_t)r * 0x00040004u * (KIND == 0 || KIND == 4 || KIND == 7 ? 0u : 1u);
        uint32_t diag = diag_next;
        diag_next = left;
        if (KIND == 0 || KIND == 7) {
            float p1l[2], p1h[2], p2l[2], p2h[2];
#pragma unroll
            for (int q = 0; q < 2; ++q) p1l[q] = p1h[q] = p2l[q] = p2h[q] = 8388608.0f;
#pragma unroll
            for (int q = 0; q < TW; ++q) {
                const uint32_t sub = prmt(tt.x, tt.y, sel[q]);
                const uint32_t up = H[q];
                bool dl, dh, ul, uh;
                const uint32_t t = __vibmax_s16x2(up, left, &uh, &ul);
                const uint32_t d = __viaddmax_s16x2(diag, sub, NEG2);
                const uint32_t h = __vibmax_s16x2(d, t, &dh, &dl);
                const float bit = (float)(1u << (q & 15));
                if (KIND == 0) {
                    if (dl) p1l[q >> 4] += bit;
                    if (dh) p1h[q >> 4] += bit;
                    if (ul) p2l[q >> 4] += bit;
                    if (uh) p2h[q >> 4] += bit;
                } else {  // the same exact sums as fused multiply-adds: 1.0 (opaque) * bit + acc
                    if (dl) p1l[q >> 4] = __fmaf_rn(one, bit, p1l[q >> 4]);
                    if (dh) p1h[q >> 4] = __fmaf_rn(one, bit, p1h[q >> 4]);
                    if (ul) p2l[q >> 4] = __fmaf_rn(one, bit, p2l[q >> 4]);
                    if (uh) p2h[q >> 4] = __fmaf_rn(one, bit, p2h[q >> 4]);
                }
                left = h;
                H[q] = h;
                diag = up;
            }
            uint4 w;
            w.x = __byte_perm(__float_as_uint(p1l[0]), __float_as_uint(p1h[0]), 0x5410);
            w.y = __byte_perm(__float_as_uint(p2l[0]), __float_as_uint(p2h[0]), 0x5410);
            w.z = __byte_perm(__float_as_uint(p1l[1]), __float_as_uint(p1h[1]), 0x5410);
            w.w = __byte_perm(__float_as_uint(p2l[1]), __float_as_uint(p2h[1]), 0x5410);
            *dp = w;
        } else if (KIND == 4) {
#pragma unroll
            for (int q = 0; q < TW; ++q) {
                const uint32_t sub = prmt(tt.x, tt.y, sel[q]);
                const uint32_t up = H[q];
                const uint32_t h = __viaddmax_s16x2(diag, sub, __vmaxs2(up, left));
                left = h;
                H[q] = h;
                diag = up;
            }
            if (r == rows - 1) *dp = make_uint4(H[0], H[1], H[2], H[29]);
        } else {
            uint32_t a1[4] = {0, 0, 0, 0}, a2[4] = {0, 0, 0, 0};
#pragma unroll
            for (int q = 0; q < TW; ++q) {
                const uint32_t sub = prmt(tt.x, tt.y, sel[q]);   // 4 s' + 2
                const uint32_t up = H[q];
                const uint32_t t = __viaddmax_s16x2(up, 0x00010001u, left);  // max(up + 1, left)
                const uint32_t h = __viaddmax_s16x2(diag, sub, t);          // max(diag + 4s' + 2, t)
                const uint32_t x = h & 0xFFFCFFFCu;
                if (KIND == 1) {
                    a1[q >> 3] = a1[q >> 3] * 4u + h;
                    a2[q >> 3] = a2[q >> 3] * 4u + x;
                } else if (KIND == 5) {  // the same, the multiply-adds spelled out
                    asm("mad.lo.u32 %0, %0, 4, %1;" : "+r"(a1[q >> 3]) : "r"(h));
                    asm("mad.lo.u32 %0, %0, 4, %1;" : "+r"(a2[q >> 3]) : "r"(x));
                } else if (KIND == 6) {  // one Horner chain per strip half: acc = acc*4 + (h - x)
                    uint32_t mtag;
                    asm("mad.lo.u32 %0, %1, -1, %2;" : "=r"(mtag) : "r"(x), "r"(h));
                    asm("mad.lo.u32 %0, %0, 4, %1;" : "+r"(a1[q >> 3]) : "r"(mtag));
                } else if (KIND == 2) {
                    a1[q >> 3] = (a1[q >> 3] << 2) | (h & 0x00030003u);
                } else {
                    a1[q >> 3] ^= h;  // (1 LOP3 so that h stays live; KIND 3 is a ceiling, not a candidate)
                }
                left = x;
                H[q] = x;
                diag = up;
            }
            *dp = make_uint4(a1[0] - a2[0], a1[1] - a2[1], a1[2] - a2[2], a1[3] - a2[3]);
        }
        dp += nthreads;
        if ((r & 7) == 7) dp = out + tid;  // stay inside a small footprint: this is not a bandwidth test
    }
    const unsigned long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int KIND>
void run(const char *name) {
    const int blocks = 148 * 4 * 4, rows = 2000;
    uint4 *out; unsigned long long *cyc, h; uint32_t *sel, hs[30 * 128];
    for (int i = 0; i < 30 * 128; ++i) hs[i] = (uint32_t)((i * 2654435761u) >> 13);
    cudaMalloc(&sel, sizeof(hs)); cudaMemcpy(sel, hs, sizeof(hs), cudaMemcpyHostToDevice);
    cudaMalloc(&out, (size_t)blocks * 128 * 8 * sizeof(uint4)); cudaMalloc(&cyc, 8);
    k<KIND><<<blocks, 128>>>(20, 1u, 1.0f, sel, out, cyc);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a);
    k<KIND><<<blocks, 128>>>(rows, 1u, 1.0f, sel, out, cyc);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    const double cellpairs = (double)blocks * 128 * rows * 30;
    // cycles one SM sub-partition spends per cell-pair of one warp: 148*4 SMSPs, 32 threads per warp instruction
    const double cyc_per = ms * 1e-3 * 1.965e9 * 148 * 4 / (cellpairs / 32);
    printf("%-44s %8.3f ms  %7.1f GCUPS-equivalent  %.2f clk per cell-pair and SMSP  (cudaError %d)\n", name, ms, 2 * cellpairs / ms * 1e-6, cyc_per,
           (int)cudaGetLastError());
    cudaFree(out); cudaFree(cyc);
}

int main() {
    run<0>("0 planes by predicates (shipped)");
    run<1>("1 in-band tags, banked by 2 IMAD");
    run<2>("2 in-band tags, banked on the integer pipe");
    run<3>("3 in-band tags, 1 LOP3 instead of banking");
    run<4>("4 score form");
    run<7>("7 planes by predicates, FFMA instead of FADD");
    run<5>("5 in-band tags, 2 mad.lo (asm)");
    run<6>("6 in-band tags, h - x then one Horner mad.lo (asm)");
    return 0;
}
