// Microbenchmark (evidence for DESIGN 3.1): what HBM3e sustains for whole 512-byte rows at RANDOM addresses, next to a contiguous
// copy -- the access pattern of the training step (K3 / K4) versus the pattern MEASURED_PEAKS.json was taken with.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a scripts/rowgather_bench.cu -o /tmp/rowgather && /tmp/rowgather
// A warp moves UNROLL rows per iteration (all loads issued before the first store), 2048 threads per SM resident.
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <vector>
#include <algorithm>
#include <random>

#define UNROLL 4
__global__ void __launch_bounds__(256) copy_rows(const float4* __restrict__ src, float4* __restrict__ dst, const int32_t* __restrict__ idx, int64_t n,
                                                 int mode) {
    // mode 0: dst[k] = src[k] (contiguous);  1: dst[k] = src[idx[k]] (random gather);  2: dst[idx[k]] = src[idx[k]] * 1.0001 (random RMW);
    // 3: dst[idx[k]] = src[k] (random scatter)
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t k0 = warp * UNROLL; k0 < n; k0 += n_warps * UNROLL) {
        float4 v[UNROLL];
        int64_t r[UNROLL];
#pragma unroll
        for (int q = 0; q < UNROLL; ++q) {
            const int64_t k = k0 + q < n ? k0 + q : n - 1;
            r[q] = mode == 0 ? k : idx[k];
            v[q] = src[(mode == 3 ? k : r[q]) * 32 + lane];
        }
#pragma unroll
        for (int q = 0; q < UNROLL; ++q) {
            if (k0 + q >= n) break;
            if (mode == 2) { v[q].x *= 1.0001f; v[q].y *= 1.0001f; v[q].z *= 1.0001f; v[q].w *= 1.0001f; }
            dst[(mode == 1 || mode == 0 ? k0 + q : r[q]) * 32 + lane] = v[q];
        }
    }
}

int main() {
    const int64_t rows = 20000000, n = 1 << 23;   // 10.2 GB table of 512-byte rows, 8M rows moved per launch (4.3 GB each way)
    float4 *table, *out;
    int32_t* idx;
    cudaMalloc(&table, rows * 512);
    cudaMalloc(&out, n * 512);
    cudaMalloc(&idx, n * 4);
    cudaMemset(table, 0, rows * 512);
    std::vector<int32_t> h(rows);
    for (int64_t k = 0; k < rows; ++k) h[k] = (int32_t)k;
    std::mt19937_64 g(1);
    std::shuffle(h.begin(), h.end(), g);
    cudaMemcpy(idx, h.data(), n * 4, cudaMemcpyHostToDevice);   // n DISTINCT random rows
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const char* names[4] = {"contiguous copy", "random 512B-row gather -> contiguous", "random 512B-row read-modify-write in place", "contiguous -> random 512B-row scatter"};
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    printf("{");
    for (int mode = 0; mode < 4; ++mode) {
        float best = 1e9f;
        for (int rep = 0; rep < 6; ++rep) {
            cudaEventRecord(e0);
            copy_rows<<<sms * 8, 256>>>(mode == 3 ? out : table, (mode == 2 || mode == 3) ? table : out, idx, n, mode);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            if (rep && ms < best) best = ms;
        }
        printf("%s\"%s\": {\"ms\": %.4f, \"GBs\": %.1f}", mode ? ", " : "", names[mode], best, 2.0 * n * 512 / best / 1e6);
    }
    printf("}\n");
    return cudaGetLastError() != cudaSuccess;
}
