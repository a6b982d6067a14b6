// Divergent-gather microbenchmark: every lane reads a random, naturally aligned record of
// 4/8/16/32 bytes from a table of `bytes` size; reports achieved records/clk/SM and GB/s.
// Answers: what does one lane's fetch cost in the L1/L2 path on B200, per access width?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gather gather.cu && ./gather
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>

struct __align__(32) F8 { float4 a, b; };
__device__ __forceinline__ F8 ldg256(const void* p) {
    F8 r;
    asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(r.a.x), "=f"(r.a.y), "=f"(r.a.z), "=f"(r.a.w), "=f"(r.b.x), "=f"(r.b.y), "=f"(r.b.z), "=f"(r.b.w) : "l"(p));
    return r;
}
__device__ __forceinline__ unsigned next(unsigned& s) { s ^= s << 13; s ^= s >> 17; s ^= s << 5; return s; }

template <int W, int UNROLL>
__global__ void __launch_bounds__(256) k_gather(const char* __restrict__ table, unsigned mask, int iters, float* out, int coherent) {
    unsigned s = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 12345u;
    float acc = 0.f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            unsigned r = next(s);
            if (coherent) r = __shfl_sync(0xffffffffu, r, 0) + (threadIdx.x & 31) * W;  // one base per warp, lanes contiguous
            const char* p = table + ((r & mask) & ~(unsigned)(W - 1));
            if (W == 4) acc += __ldg((const float*)p);
            else if (W == 8) { float2 v = __ldg((const float2*)p); acc += v.x + v.y; }
            else if (W == 16) { float4 v = __ldg((const float4*)p); acc += v.x + v.y + v.z + v.w; }
            else { F8 v = ldg256(p); acc += v.a.x + v.a.y + v.a.z + v.a.w + v.b.x + v.b.y + v.b.z + v.b.w; }
        }
    }
    if (acc == 123.456f) out[0] = acc;
}

template <int W>
void run(const char* d, size_t bytes, int coherent) {
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
    float* out; cudaMalloc(&out, 4);
    const int iters = 2000, U = 4, grid = prop.multiProcessorCount * 8, block = 256;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k_gather<W, U><<<grid, block>>>(d, (unsigned)(bytes - 1), 100, out, coherent);
    cudaEventRecord(e0);
    k_gather<W, U><<<grid, block>>>(d, (unsigned)(bytes - 1), iters, out, coherent);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double recs = (double)grid * block * iters * U;
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    double cyc = ms * 1e-3 * clk * 1e3;
    printf("table %8zu KB  W=%2d B %s: %7.3f ms  %8.1f Grec/s  %8.1f GB/s  %.3f rec/clk/SM (at %d MHz nominal)\n", bytes >> 10, W,
           coherent ? "coalesced" : "divergent", ms, recs / ms / 1e6, recs * W / ms / 1e6, recs / cyc / prop.multiProcessorCount, clk / 1000);
    cudaFree(out);
}

int main() {
    size_t sizes[] = {32u << 10, 128u << 10, 1u << 20, 8u << 20, 64u << 20};
    for (size_t bytes : sizes) {
        char* d; cudaMalloc(&d, bytes); cudaMemset(d, 0, bytes);
        for (int coherent = 0; coherent < 2; ++coherent) {
            run<4>(d, bytes, coherent); run<8>(d, bytes, coherent); run<16>(d, bytes, coherent); run<32>(d, bytes, coherent);
        }
        cudaFree(d);
    }
    return 0;
}
