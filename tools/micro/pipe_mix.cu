// pipe_mix.cu -- issue-rate microbenchmark: which instruction mixes share an issue cadence on sm_100a?
// Every kernel runs CH independent dependent-chains per thread of the named instruction mix; the result is
// warp-instructions per cycle per SM sub-partition (SMSP), measured with clock64 on a full grid.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_mix pipe_mix.cu ; run: ./pipe_mix
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t s) { uint32_t r; asm volatile("prmt.b32 %0,%1,%2,%3;" : "=r"(r) : "r"(a), "r"(b), "r"(s)); return r; }
__device__ __forceinline__ float fadd_imm(float a) { float r; asm volatile("add.f32 %0,%1,0f40800000;" : "=f"(r) : "f"(a)); return r; }
__device__ __forceinline__ float fadd_pred(float a, bool p) { float r = a; asm volatile("{.reg .pred q; setp.ne.u32 q,%1,0; @q add.f32 %0,%0,0f40800000;}" : "+f"(r) : "r"((uint32_t)p)); return r; }
__device__ __forceinline__ float ffma_reg(float a, float b, float c) { float r; asm volatile("fma.rn.f32 %0,%1,%2,%3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }
__device__ __forceinline__ uint32_t iadd3(uint32_t a, uint32_t b) { uint32_t r; asm volatile("add.u32 %0,%1,%2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ uint32_t imad(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("mad.lo.u32 %0,%1,%2,%3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }

template <int KIND>
__global__ void __launch_bounds__(256) k(int iters, uint32_t seed, uint32_t one, uint32_t *sink, unsigned long long *cyc) {
    constexpr int CH = 8;
    uint32_t v[CH], u[CH];
    float f[CH];
    const uint32_t g = 0xFFFDFFFDu ^ (seed & 1);
    uint32_t w = seed * 2654435761u + threadIdx.x;
#pragma unroll
    for (int c = 0; c < CH; ++c) { v[c] = (threadIdx.x + c * 7 + seed) & 0x00FF00FF; u[c] = v[c] ^ 0x5; f[c] = 8388608.0f; }
    const unsigned long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
#pragma unroll
            for (int c = 0; c < CH; ++c) {
                if (KIND == 0) { v[c] = __viaddmax_s16x2(v[c], g, w); }                                   // ALU only
                if (KIND == 1) { v[c] = __viaddmax_s16x2(v[c], g, w); u[c] = imad(u[c], one, w); }        // ALU + IMAD reg
                if (KIND == 2) { u[c] = imad(u[c], one, w); }                                             // IMAD only
                if (KIND == 3) { v[c] = __viaddmax_s16x2(v[c], g, w); f[c] = fadd_imm(f[c]); }                     // ALU + FADD imm
                if (KIND == 4) { f[c] = fadd_imm(f[c]); }                                                          // FADD imm only
                if (KIND == 5) { v[c] = __viaddmax_s16x2(v[c], g, w); u[c] = prmt(u[c], w, v[c]); }       // ALU + ALU
                if (KIND == 6) { bool ph, pl; v[c] = __vibmax_s16x2(v[c], w, &ph, &pl); if (pl) f[c] += 4.0f; if (ph) f[c] += 2.0f; }  // VIMNMX pred + 2 FADD
                if (KIND == 7) { v[c] = __viaddmax_s16x2(v[c], g, w); f[c] = ffma_reg(f[c], __uint_as_float(one), __uint_as_float(w)); }  // ALU + FFMA reg
                if (KIND == 8) { v[c] = __viaddmax_s16x2(v[c], g, w); f[c] = fadd_imm(f[c]); f[c] = fadd_imm(f[c]); }       // ALU + 2 FADD imm
                if (KIND == 9) { v[c] = __viaddmax_s16x2(v[c], g, w); u[c] = imad(u[c], one, w); f[c] = fadd_imm(f[c]); f[c] = fadd_imm(f[c]); }  // ALU + IMAD + 2 FADD imm
                if (KIND == 10) { v[c] = __viaddmax_s16x2(v[c], g, w); u[c] = iadd3(u[c], w); }          // ALU + IADD3
                if (KIND == 12) { v[c] = __viaddmax_s16x2(v[c], g, w); f[c] = ffma_reg(__uint_as_float(one), 4.0f, f[c]); }   // ALU + FFMA imm
                if (KIND == 13) { v[c] = __viaddmax_s16x2(v[c], g, w); f[c] = ffma_reg(__uint_as_float(one), 4.0f, f[c]); f[c] = ffma_reg(__uint_as_float(one), 2.0f, f[c]); }   // ALU + 2 FFMA imm
                if (KIND == 14) { bool ph, pl; v[c] = __vibmax_s16x2(v[c], w, &ph, &pl); if (pl) f[c] = ffma_reg(__uint_as_float(one), 4.0f, f[c]); if (ph) f[c] = ffma_reg(__uint_as_float(one), 2.0f, f[c]); }  // VIMNMX pred + 2 @p FFMA
                if (KIND == 15) { v[c] = __viaddmax_s16x2(v[c], g, w); f[c] = ffma_reg(f[c], __uint_as_float(one), __uint_as_float(w)); f[c] = ffma_reg(f[c], __uint_as_float(one), __uint_as_float(w));}  // ALU + 2 FFMA reg
                if (KIND == 16) { f[c] = ffma_reg(f[c], __uint_as_float(one), __uint_as_float(w)); }  // FFMA reg only
                if (KIND == 17) { v[c] = __viaddmax_s16x2(v[c], g, w); float t; asm volatile("mul.f32 %0,%1,%2;" : "=f"(t) : "f"(f[c]), "f"(__uint_as_float(one))); f[c] = t; }  // ALU + FMUL
                if (KIND == 18) { v[c] = __viaddmax_s16x2(v[c], g, w); float t; asm volatile("add.f32 %0,%1,%2;" : "=f"(t) : "f"(f[c]), "f"(__uint_as_float(w))); f[c] = t; }  // ALU + FADD reg
                if (KIND == 19) { bool ph, pl; v[c] = __vibmax_s16x2(v[c], w, &ph, &pl); if (pl) f[c] = ffma_reg(__uint_as_float(one), __uint_as_float(g), f[c]); if (ph) f[c] = ffma_reg(__uint_as_float(one), __uint_as_float(w), f[c]); }  // VIMNMX pred + 2 @p FFMA reg
                if (KIND == 11) { v[c] = __viaddmax_s16x2(v[c], g, w); u[c] = imad(u[c], 3u, w); }        // ALU + IMAD imm-multiplier
            }
            w += 0x00010001u;
        }
    }
    const unsigned long long t1 = clock64();
    uint32_t acc = 0;
#pragma unroll
    for (int c = 0; c < CH; ++c) acc ^= v[c] ^ u[c] ^ __float_as_uint(f[c]);
    if (acc == 0x12345678u) sink[0] = acc;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int KIND>
void run(const char *name, int per_iter_instr) {
    uint32_t *sink; unsigned long long *cyc, h;
    cudaMalloc(&sink, 64); cudaMalloc(&cyc, 8);
    const int iters = 2000, blocks = 148 * 8;
    k<KIND><<<blocks, 256>>>(10, 1u, 1u, sink, cyc);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a);
    k<KIND><<<blocks, 256>>>(iters, 1u, 1u, sink, cyc);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    // per SMSP: 8 blocks * 8 warps / 4 = 16 warps; instr per warp = iters*8*8*per_iter_instr
    const double instr_smsp = 16.0 * iters * 64.0 * per_iter_instr;
    const double cycles = ms * 1e-3 * 1.965e9;  // assumes max clock; clock64 of block 0 printed beside it
    printf("%-34s %2d instr/step: %.3f warp-instr/clk/SMSP (event, @1.965GHz)  block0 clk %.0f -> %.3f\n", name, per_iter_instr,
           instr_smsp / cycles, (double)h, instr_smsp / (double)h);
    cudaFree(sink); cudaFree(cyc);
}

int main() {
    run<0>("VIADDMNMX", 1);
    run<2>("IMAD reg", 1);
    run<4>("FADD imm", 1);
    run<1>("VIADDMNMX + IMAD reg", 2);
    run<11>("VIADDMNMX + IMAD imm", 2);
    run<3>("VIADDMNMX + FADD imm", 2);
    run<8>("VIADDMNMX + 2 FADD imm", 3);
    run<9>("VIADDMNMX + IMAD + 2 FADD imm", 4);
    run<5>("VIADDMNMX + PRMT", 2);
    run<10>("VIADDMNMX + IADD3", 2);
    run<7>("VIADDMNMX + FFMA reg", 2);
    run<6>("VIMNMX.pred + 2 @p FADD imm", 3);
    run<16>("FFMA reg", 1);
    run<12>("VIADDMNMX + FFMA imm", 2);
    run<13>("VIADDMNMX + 2 FFMA imm", 3);
    run<15>("VIADDMNMX + 2 FFMA reg", 3);
    run<17>("VIADDMNMX + FMUL reg", 2);
    run<18>("VIADDMNMX + FADD reg", 2);
    run<14>("VIMNMX.pred + 2 @p FFMA imm", 3);
    run<19>("VIMNMX.pred + 2 @p FFMA reg", 3);
    return 0;
}
