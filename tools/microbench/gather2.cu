// L1 data-stage behaviour for the access pattern of a BVH node visit: every lane of a warp
// loads ROWS (16 B) of a different random record.  Variables: record stride (112 / 128 B),
// rows per visit, and how many lanes are active.  Reports lane-rows per clock per SM.
#include <cuda_runtime.h>
#include <cstdio>
__device__ __forceinline__ unsigned nextr(unsigned& s) { s ^= s << 13; s ^= s >> 17; s ^= s << 5; return s; }

template <int STRIDE, int ROWS>
__global__ void __launch_bounds__(128) k(const char* __restrict__ table, unsigned nrec, int iters, float* out, unsigned laneMask) {
    unsigned s = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 12345u;
    float acc = 0.f;
    const bool on = (laneMask >> (threadIdx.x & 31)) & 1u;
    for (int it = 0; it < iters; ++it) {
        const unsigned r = nextr(s) % nrec;
        if (on) {
            const float4* p = (const float4*)(table + (size_t)r * STRIDE);
#pragma unroll
            for (int k = 0; k < ROWS; ++k) { float4 v = __ldg(p + k); acc += v.x + v.y + v.z + v.w; }
        }
    }
    if (acc == 123.456f) out[0] = acc;
}

template <int STRIDE, int ROWS>
void run(const char* d, size_t bytes, unsigned laneMask) {
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
    float* out; cudaMalloc(&out, 4);
    const int iters = 4000, grid = prop.multiProcessorCount * 12, block = 128;
    const unsigned nrec = (unsigned)(bytes / STRIDE);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<STRIDE, ROWS><<<grid, block>>>(d, nrec, 100, out, laneMask);
    cudaEventRecord(e0);
    k<STRIDE, ROWS><<<grid, block>>>(d, nrec, iters, out, laneMask);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const int lanes = __builtin_popcount(laneMask);
    const double rows = (double)grid * (block / 32) * lanes * iters * ROWS;
    const double cyc = ms * 1e-3 * clk * 1e3;
    printf("table %6zu KB stride %3d rows %d lanes %2d: %7.3f ms  %.3f lane-rows/clk/SM  (%.3f visits/clk/SM)\n", bytes >> 10, STRIDE, ROWS, lanes, ms,
           rows / cyc / prop.multiProcessorCount, rows / ROWS / cyc / prop.multiProcessorCount);
    cudaFree(out);
}

int main() {
    size_t sizes[] = {96u << 10, 2u << 20};
    unsigned masks[] = {0xffffffffu, 0x0000ffffu, 0x000000ffu, 0x11111111u};
    for (size_t bytes : sizes) {
        char* d; cudaMalloc(&d, bytes + 256); cudaMemset(d, 0, bytes + 256);
        for (unsigned m : masks) {
            run<128, 7>(d, bytes, m); run<112, 7>(d, bytes, m); run<128, 1>(d, bytes, m); run<112, 1>(d, bytes, m);
            run<128, 4>(d, bytes, m); run<64, 4>(d, bytes, m); run<80, 5>(d, bytes, m); run<48, 3>(d, bytes, m);
        }
        cudaFree(d);
    }
    return 0;
}
