// Is the texture path an independent gather engine on B200?  Every lane fetches ROWS (float4) of a random
// 112-byte record, through LSU loads (ld.global.nc), through tex1Dfetch on a linear texture object, or a
// mix (TEXROWS of the 7 rows through the texture unit).  Reports lane-rows/clk/SM.
#include <cuda_runtime.h>
#include <cstdio>
__device__ __forceinline__ unsigned nextr(unsigned& s) { s ^= s << 13; s ^= s >> 17; s ^= s << 5; return s; }

template <int TEXROWS>
__global__ void __launch_bounds__(128) k(const float4* __restrict__ table, cudaTextureObject_t tex, unsigned nrec, int iters, float* out, unsigned laneMask) {
    unsigned s = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 12345u;
    float acc = 0.f;
    const bool on = (laneMask >> (threadIdx.x & 31)) & 1u;
    for (int it = 0; it < iters; ++it) {
        const unsigned r = nextr(s) % nrec;
        if (on) {
#pragma unroll
            for (int k = 0; k < 7; ++k) {
                float4 v;
                if (k < TEXROWS) v = tex1Dfetch<float4>(tex, (int)(r * 7 + k));
                else v = __ldg(table + (size_t)r * 7 + k);
                acc += v.x + v.y + v.z + v.w;
            }
        }
    }
    if (acc == 123.456f) out[0] = acc;
}

template <int TEXROWS>
void run(const float4* d, cudaTextureObject_t tex, size_t bytes, unsigned laneMask) {
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
    float* out; cudaMalloc(&out, 4);
    const int iters = 3000, grid = prop.multiProcessorCount * 12, block = 128;
    const unsigned nrec = (unsigned)(bytes / 112);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<TEXROWS><<<grid, block>>>(d, tex, nrec, 100, out, laneMask);
    cudaEventRecord(e0);
    k<TEXROWS><<<grid, block>>>(d, tex, nrec, iters, out, laneMask);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const int lanes = __builtin_popcount(laneMask);
    const double rows = (double)grid * (block / 32) * lanes * iters * 7;
    const double cyc = ms * 1e-3 * clk * 1e3;
    printf("table %6zu KB  tex rows %d of 7  lanes %2d: %7.3f ms  %.3f lane-rows/clk/SM  err=%s\n", bytes >> 10, TEXROWS, lanes, ms,
           rows / cyc / prop.multiProcessorCount, cudaGetErrorString(cudaGetLastError()));
    cudaFree(out);
}

int main() {
    size_t sizes[] = {96u << 10, 2u << 20};
    unsigned masks[] = {0xffffffffu, 0x0000ffffu, 0x000000ffu};
    for (size_t bytes : sizes) {
        float4* d; cudaMalloc(&d, bytes + 256); cudaMemset(d, 0, bytes + 256);
        cudaResourceDesc rd = {}; rd.resType = cudaResourceTypeLinear; rd.res.linear.devPtr = d; rd.res.linear.desc = cudaCreateChannelDesc<float4>();
        rd.res.linear.sizeInBytes = bytes;
        cudaTextureDesc td = {}; td.readMode = cudaReadModeElementType; td.filterMode = cudaFilterModePoint; td.addressMode[0] = cudaAddressModeClamp;
        cudaTextureObject_t tex = 0; cudaCreateTextureObject(&tex, &rd, &td, nullptr);
        for (unsigned m : masks) { run<0>(d, tex, bytes, m); run<7>(d, tex, bytes, m); run<3>(d, tex, bytes, m); run<2>(d, tex, bytes, m); run<4>(d, tex, bytes, m); }
        cudaDestroyTextureObject(tex); cudaFree(d);
    }
    return 0;
}
