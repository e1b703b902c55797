// gather_bench.cu -- how fast can one B200 gather random 256-byte rows (the access pattern of a
// D = 64 fp32 SpMM) as a function of the table size?  Sizes the real ceiling of K1: below the L2
// capacity it measures the L2 -> SM gather rate, above it the DRAM random-row rate.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o gather_bench gather_bench.cu
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <vector>

__global__ void __launch_bounds__(256) gather_rows(const float4* __restrict__ X, const uint32_t* __restrict__ idx,
                                                   float4* __restrict__ out, int64_t n_idx, int unroll_dummy)
{
    // 16 lanes per 256-byte row, 2 rows per warp step, 8 rows per lane-batch in flight
    const int lane = threadIdx.x & 31, sub = lane & 15, grp = lane >> 4;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t per_warp = 64;
    const int64_t base = warp * per_warp;
    if (base >= n_idx) return;
    float4 acc = make_float4(0, 0, 0, 0);
    for (int j = 0; j < per_warp; j += 16) {
        float4 v[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const uint32_t r = idx[base + j + q * 2 + grp];
            v[q] = __ldg(X + (size_t)r * 16 + sub);
        }
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            acc.x += v[q].x; acc.y += v[q].y; acc.z += v[q].z; acc.w += v[q].w;
        }
    }
    if (acc.x == 123.456f) out[warp] = acc;  // keep the loads alive
}

int main()
{
    const int64_t n_idx = 64ll << 20;  // 64M gathered rows = 16 GiB of gather traffic
    uint32_t* d_idx;
    cudaMalloc(&d_idx, n_idx * 4);
    float4* d_out;
    cudaMalloc(&d_out, (n_idx / 64) * 16);
    std::vector<uint32_t> h(n_idx);
    const size_t sizes_mb[] = {8, 16, 32, 48, 64, 96, 128, 192, 256, 384, 512, 1024, 4096};
    for (size_t mb : sizes_mb) {
        const size_t rows = mb * 1024 * 1024 / 256;
        float4* X;
        cudaMalloc(&X, rows * 256);
        cudaMemset(X, 0, rows * 256);
        uint64_t s = 88172645463325252ull;
        for (int64_t i = 0; i < n_idx; ++i) {
            s ^= s << 13; s ^= s >> 7; s ^= s << 17;
            h[i] = (uint32_t)(s % rows);
        }
        cudaMemcpy(d_idx, h.data(), n_idx * 4, cudaMemcpyHostToDevice);
        const int64_t warps = n_idx / 64;
        const int blocks = (int)((warps * 32 + 255) / 256);
        cudaEvent_t a, b;
        cudaEventCreate(&a);
        cudaEventCreate(&b);
        for (int w = 0; w < 2; ++w) gather_rows<<<blocks, 256>>>(X, d_idx, d_out, n_idx, 0);
        cudaEventRecord(a);
        const int reps = 3;
        for (int w = 0; w < reps; ++w) gather_rows<<<blocks, 256>>>(X, d_idx, d_out, n_idx, 0);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms;
        cudaEventElapsedTime(&ms, a, b);
        ms /= reps;
        printf("table %5zu MB: %.3f ms for %lld row gathers -> %.0f GB/s gathered (+%.0f GB/s index stream)\n", mb, ms,
               (long long)n_idx, n_idx * 256.0 / ms / 1e6, n_idx * 4.0 / ms / 1e6);
        cudaFree(X);
    }
    return 0;
}
