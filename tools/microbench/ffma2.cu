// Issue rate of Blackwell's packed FP32 instructions (FFMA2 / FADD2) against scalar FFMA, alone and mixed with integer
// (ALU-pipe) instructions: does a packed instruction free an issue slot, or does it hold the scheduler for two cycles?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2 ffma2.cu && ./ffma2
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ float fma1(float a, float b, float c) { float r; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }
__device__ __forceinline__ unsigned lop(unsigned a, unsigned b) { unsigned r; asm volatile("xor.b32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }

template <int MODE>  // 0: 8 FFMA  1: 8 FFMA2  2: 8 FFMA + 8 LOP  3: 8 FFMA2 + 8 LOP  4: 16 FFMA + 8 LOP   5: 8 LOP
__global__ void __launch_bounds__(1024, 1) k(float* out, int iters, float s) {
    float f[16]; u64 p[8]; unsigned q[8];
    for (int i = 0; i < 16; ++i) f[i] = s + i + threadIdx.x;
    for (int i = 0; i < 8; ++i) { p[i] = ((u64)__float_as_uint(s + i) << 32) | __float_as_uint(s + threadIdx.x); q[i] = threadIdx.x + i; }
    u64 ps = ((u64)__float_as_uint(s) << 32) | __float_as_uint(s);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0 || MODE == 2 || MODE == 4) f[i] = fma1(f[i], s, s);
            if (MODE == 4) f[8 + i] = fma1(f[8 + i], s, s);
            if (MODE == 1 || MODE == 3) p[i] = fma2(p[i], ps, ps);
            if (MODE == 2 || MODE == 3 || MODE == 4 || MODE == 5) q[i] = lop(q[i], q[(i + 1) & 7]);
        }
    }
    float acc = 0; for (int i = 0; i < 16; ++i) acc += f[i];
    for (int i = 0; i < 8; ++i) acc += __uint_as_float((unsigned)p[i]) + __uint_as_float((unsigned)(p[i] >> 32)) + q[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
template <int MODE> void run(const char* name, int perIter) {
    float* out; cudaMalloc(&out, 148 * 1024 * 4);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    const int iters = 200000;
    k<MODE><<<148, 1024>>>(out, 1000, 1.0001f);
    cudaEventRecord(a); k<MODE><<<148, 1024>>>(out, iters, 1.0001f); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const double warpInstr = 32.0 * iters * perIter;            // per SM: 32 warps
    const double cycles = ms * 1e-3 * clk * 1e3;
    printf("%-28s %8.3f ms  %6.3f warp-instr/clk/SM (of 4)   %.2f cycles per loop body per scheduler\n", name, ms, warpInstr / cycles, cycles / iters / 8.0);
    cudaFree(out);
}
int main() {
    run<0>("8 FFMA", 8); run<1>("8 FFMA2", 8); run<5>("8 LOP", 8); run<2>("8 FFMA + 8 LOP", 16); run<3>("8 FFMA2 + 8 LOP", 16); run<4>("16 FFMA + 8 LOP", 24);
    return 0;
}
