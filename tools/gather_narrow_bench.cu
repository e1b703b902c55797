// gather_narrow_bench.cu -- ceiling of the column-sharded SpMM's gathers: random rows of 32 / 64 / 128 / 256 bytes from
// an L2-resident table (a rank of an 8-GPU column-sharded run gathers 32-byte rows from a 48 MB table), 128-bit loads,
// 8 in flight per lane, perfect index streaming, no arithmetic.  Reports sectors/s and bytes/s.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o gather_narrow_bench gather_narrow_bench.cu
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <vector>

template <int LPR>  // lanes (x 16 bytes) per row
__global__ void __launch_bounds__(256) gather_rows(const float4* __restrict__ X, const uint32_t* __restrict__ idx,
                                                   float4* __restrict__ out, int64_t n_idx)
{
    constexpr int RPW = 32 / LPR;  // rows per warp step
    const int lane = threadIdx.x & 31, sub = lane % LPR, grp = lane / LPR;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t per_warp = 64 * RPW;
    const int64_t base = warp * per_warp;
    if (base >= n_idx) return;
    float4 acc = make_float4(0, 0, 0, 0);
    for (int j = 0; j < per_warp; j += 8 * RPW) {
        float4 v[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const uint32_t r = idx[base + j + q * RPW + grp];
            v[q] = __ldg(X + (size_t)r * LPR + sub);
        }
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            acc.x += v[q].x; acc.y += v[q].y; acc.z += v[q].z; acc.w += v[q].w;
        }
    }
    if (acc.x == 123.456f) out[warp] = acc;
}

template <int LPR>
static void run(size_t mb, int64_t n_idx, uint32_t* d_idx, std::vector<uint32_t>& h, float4* d_out)
{
    const size_t row_bytes = 16 * LPR;
    const size_t rows = mb * 1024 * 1024 / row_bytes;
    float4* X;
    cudaMalloc(&X, rows * row_bytes);
    cudaMemset(X, 0, rows * row_bytes);
    uint64_t s = 88172645463325252ull;
    for (int64_t i = 0; i < n_idx; ++i) {
        s ^= s << 13; s ^= s >> 7; s ^= s << 17;
        h[i] = (uint32_t)(s % rows);
    }
    cudaMemcpy(d_idx, h.data(), n_idx * 4, cudaMemcpyHostToDevice);
    const int64_t warps = n_idx / (64 * (32 / LPR));
    const int blocks = (int)((warps * 32 + 255) / 256);
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    float best = 1e9f;
    for (int it = 0; it < 5; ++it) {
        cudaEventRecord(a);
        gather_rows<LPR><<<blocks, 256>>>(X, d_idx, d_out, n_idx);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms;
        cudaEventElapsedTime(&ms, a, b);
        if (it > 0 && ms < best) best = ms;
    }
    printf("row %3zu B  table %4zu MB: %.3f ms for %lld row gathers -> %.1f G rows/s, %.1f G sectors/s, %.0f GB/s gathered\n",
           row_bytes, mb, best, (long long)n_idx, n_idx / best * 1e-6, n_idx * (row_bytes / 32.0) / best * 1e-6,
           n_idx * (double)row_bytes / best * 1e-6);
    cudaFree(X);
}

int main()
{
    const int64_t n_idx = 128ll << 20;
    uint32_t* d_idx;
    cudaMalloc(&d_idx, n_idx * 4);
    float4* d_out;
    cudaMalloc(&d_out, (n_idx / 64) * 16);
    std::vector<uint32_t> h(n_idx);
    for (size_t mb : {16, 48, 96}) {
        run<2>(mb, n_idx, d_idx, h, d_out);
        run<4>(mb, n_idx, d_idx, h, d_out);
        run<8>(mb, n_idx, d_idx, h, d_out);
        run<16>(mb, n_idx / 2, d_idx, h, d_out);
    }
    return 0;
}
