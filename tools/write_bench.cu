// Microbenchmark: HBM write stream driven by 1-D bulk async stores from ONE constant shared-memory buffer, optionally
// followed (one band behind) by scattered red.global.max/min on the freshly written lines (L2 hits expected).
// usage: write_bench <total_MB> <store_bytes> <stores_per_band> <ctas_per_sm> <threads> <reds_per_band> <lag>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bulk_s2g(void* dst, const void* src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src)), "r"(bytes) : "memory");
}

__global__ void __launch_bounds__(1024, 1) write_kernel(float* dst, size_t total_bytes, int store_bytes, int spb, int reds, int lag) {
    extern __shared__ __align__(128) unsigned char smem[];
    float* src = reinterpret_cast<float*>(smem);
    const int tid = threadIdx.x, nthr = blockDim.x;
    for (int i = tid; i < store_bytes / 4; i += nthr) src[i] = (i % 15 == 14) ? 1.0f : 0.0f;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    const size_t band_bytes = (size_t)store_bytes * spb;
    const size_t n_band = total_bytes / band_bytes;
    const size_t b0 = blockIdx.x * n_band / gridDim.x, b1 = (blockIdx.x + 1) * n_band / gridDim.x;
    const int n = (int)(b1 - b0);
    for (int i = 0; i < n + lag; ++i) {
        if (tid == 0) {
            if (i < n) {
                unsigned char* d = reinterpret_cast<unsigned char*>(dst) + (b0 + i) * band_bytes;
                for (int k = 0; k < spb; ++k) bulk_s2g(d + (size_t)k * store_bytes, src, store_bytes);
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
            if (reds) {   // band i - lag must be complete before it is modified
                if (i < n) {
                    if (lag == 1) asm volatile("cp.async.bulk.wait_group 1;" ::: "memory");
                    else if (lag == 2) asm volatile("cp.async.bulk.wait_group 2;" ::: "memory");
                    else asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
                } else {
                    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
                }
                asm volatile("fence.proxy.async;" ::: "memory");
            }
        }
        if (reds) {
            __syncthreads();
            const int j = i - lag;
            if (j >= 0 && j < n) {
                int* band = reinterpret_cast<int*>(reinterpret_cast<unsigned char*>(dst) + (b0 + j) * band_bytes);
                const int band_px = (int)(band_bytes / 60);
                // runs of 24 consecutive pixels (like an object row), channel 3 max and channel 14 min
                for (int r = tid; r < reds; r += nthr) {
                    const int run = r / 24, off = r % 24;
                    const int px = (int)(((unsigned)run * 2654435761u) % (unsigned)(band_px - 24)) + off;
                    atomicMax(band + px * 15 + 3, __float_as_int(0.5f));
                    atomicMin(band + px * 15 + 14, __float_as_int(0.25f));
                }
            }
        }
    }
    if (tid == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

int main(int argc, char** argv) {
    const size_t total = (size_t)atoi(argv[1]) << 20;
    const int store_bytes = atoi(argv[2]), spb = atoi(argv[3]), cps = atoi(argv[4]), threads = atoi(argv[5]), reds = atoi(argv[6]),
              lag = atoi(argv[7]);
    float *a, *b;
    cudaMalloc(&a, total);
    cudaMalloc(&b, total);
    cudaFuncSetAttribute(write_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, store_bytes);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (int i = 0; i < 3; ++i) write_kernel<<<148 * cps, threads, store_bytes>>>(i & 1 ? a : b, total, store_bytes, spb, reds, lag);
    cudaDeviceSynchronize();
    const int n = 10;
    cudaEventRecord(e0);
    for (int i = 0; i < n; ++i) write_kernel<<<148 * cps, threads, store_bytes>>>(i & 1 ? a : b, total, store_bytes, spb, reds, lag);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    ms /= n;
    printf("%s: store %d B x %d per band, %d CTA/SM, %d thr, %d reds/band, lag %d: %.4f ms  %.0f GB/s  (%s)\n", argv[0], store_bytes, spb, cps,
           threads, reds, lag, ms, total / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
    // check
    float h[30];
    cudaMemcpy(h, a, sizeof(h), cudaMemcpyDeviceToHost);
    printf("  first pixel: %g %g ... %g | %g\n", h[0], h[3], h[14], h[29]);
    return 0;
}
