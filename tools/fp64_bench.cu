// Microbenchmark: latency (dependent chain, 1 warp) and throughput (16 warps/SM, independent chains) of DFMA, DMUL and
// F2F.F32.F64 on this GPU, in SM cycles per warp instruction.
#include <cstdio>
#include <cuda_runtime.h>

template <int OP>
__global__ void k(double* out, long long* cyc, int iters) {
    double a = threadIdx.x * 1e-3 + 1.0, b = 1.0000001, c = 1e-9;
    double a2 = a + 1, a3 = a + 2, a4 = a + 3;
    float f = 0.f, f2 = 0.f, f3 = 0.f, f4 = 0.f;
    __syncthreads();
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        if (OP == 0) { a = fma(a, b, c); }                                  // dependent DFMA chain
        if (OP == 1) { a = fma(a, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c); a4 = fma(a4, b, c); }   // 4 independent
        if (OP == 2) { f += (float)a; a += 1.0; }                            // F2F + DADD chain
        if (OP == 3) { f += (float)a; f2 += (float)a2; f3 += (float)a3; f4 += (float)a4; a = a * b; a2 = a2 * b; a3 = a3 * b; a4 = a4 * b; }
    }
    const long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = a + a2 + a3 + a4 + f + f2 + f3 + f4;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

int main() {
    double* out;
    long long* cyc;
    cudaMalloc(&out, 148 * 1024 * 8);
    cudaMallocManaged(&cyc, 8);
    const int iters = 4096;
    const char* names[] = {"DFMA dependent", "DFMA x4 independent", "F2F+DADD dependent", "4x(F2F + DMUL) independent"};
    const int per_iter[] = {1, 4, 2, 8};
    for (int threads : {32, 512}) {
        for (int op = 0; op < 4; ++op) {
            for (int w = 0; w < 2; ++w) {
                if (op == 0) k<0><<<148, threads>>>(out, cyc, iters);
                if (op == 1) k<1><<<148, threads>>>(out, cyc, iters);
                if (op == 2) k<2><<<148, threads>>>(out, cyc, iters);
                if (op == 3) k<3><<<148, threads>>>(out, cyc, iters);
                cudaDeviceSynchronize();
            }
            printf("%4d threads/SM  %-32s %8.2f cycles per iteration (%d fp64-pipe instr each) -> %.2f cycles per warp-instr per SM\n", threads,
                   names[op], (double)*cyc / iters, per_iter[op], (double)*cyc / iters / per_iter[op] / (threads / 32));
        }
    }
    return 0;
}
