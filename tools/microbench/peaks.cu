// FP32 FFMA issue peak and L2 read bandwidth of this B200 (the roofline denominators next to
// MEASURED_PEAKS.json's HBM and bf16 figures).  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o peaks peaks.cu
#include <cuda_runtime.h>
#include <cstdio>

__global__ void __launch_bounds__(256) k_ffma(float* out, int iters) {
    float a0 = threadIdx.x * 1e-3f, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const float b = 1.000001f, c = 1e-7f;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            a0 = fmaf(a0, b, c); a1 = fmaf(a1, b, c); a2 = fmaf(a2, b, c); a3 = fmaf(a3, b, c);
            a4 = fmaf(a4, b, c); a5 = fmaf(a5, b, c); a6 = fmaf(a6, b, c); a7 = fmaf(a7, b, c);
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

__global__ void __launch_bounds__(256) k_read(const float4* __restrict__ buf, size_t n4, int reps, float* out) {
    float acc = 0.f;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (int r = 0; r < reps; ++r)
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
            const float4 v = __ldcg(buf + i);  // L2 only (bypass L1)
            acc += v.x + v.y + v.z + v.w;
        }
    if (acc == 123.456f) out[0] = acc;
}

int main() {
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    float* out; cudaMalloc(&out, (size_t)prop.multiProcessorCount * 8 * 256 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms;
    {   // FFMA: 8 independent chains per thread, 8 CTAs x 256 threads per SM
        const int iters = 20000, grid = prop.multiProcessorCount * 8;
        k_ffma<<<grid, 256>>>(out, 100);
        cudaEventRecord(e0); k_ffma<<<grid, 256>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        const double flop = 2.0 * 8 * 16 * (double)iters * grid * 256;
        printf("FFMA: %.3f ms -> %.2f TFLOP/s measured (nominal %d SM x 128 lanes x 2 x %d MHz = %.2f TFLOP/s)\n", ms, flop / ms / 1e9,
               prop.multiProcessorCount, clk / 1000, prop.multiProcessorCount * 128 * 2.0 * clk * 1e3 / 1e12);
    }
    for (size_t mb : {8, 32, 64, 96}) {  // L2-resident working sets (126 MB L2)
        const size_t bytes = mb << 20, n4 = bytes / 16;
        float4* buf; cudaMalloc(&buf, bytes); cudaMemset(buf, 0, bytes);
        const int grid = prop.multiProcessorCount * 8, reps = 40;
        k_read<<<grid, 256>>>(buf, n4, 2, out);
        cudaEventRecord(e0); k_read<<<grid, 256>>>(buf, n4, reps, out); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        printf("L2 read, %3zu MB working set: %.3f ms -> %.0f GB/s\n", mb, ms, (double)bytes * reps / ms / 1e6);
        cudaFree(buf);
    }
    printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
