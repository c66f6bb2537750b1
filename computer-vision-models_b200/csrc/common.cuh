// Shared device/host helpers for libcvmhot (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "cvmhot.h"

#define CVM_NUM_SMS_FALLBACK 148

// ---- host-side error plumbing --------------------------------------------------------------------------------------
void cvm_set_error(const char* fmt, ...);
int cvm_num_sms();

#define CVM_CHECK_ARG(cond, ...)                 \
    do {                                         \
        if (!(cond)) {                           \
            cvm_set_error(__VA_ARGS__);          \
            return CVM_ERR_ARG;                  \
        }                                        \
    } while (0)

#define CVM_CHECK_CUDA(expr)                                                            \
    do {                                                                                \
        cudaError_t _e = (expr);                                                        \
        if (_e != cudaSuccess) {                                                        \
            cvm_set_error("%s failed: %s", #expr, cudaGetErrorString(_e));              \
            return CVM_ERR_CUDA;                                                        \
        }                                                                               \
    } while (0)

#define CVM_CHECK_LAUNCH(name)                                                          \
    do {                                                                                \
        cudaError_t _e = cudaGetLastError();                                            \
        if (_e != cudaSuccess) {                                                        \
            cvm_set_error("launch of %s failed: %s", name, cudaGetErrorString(_e));     \
            return CVM_ERR_CUDA;                                                        \
        }                                                                               \
    } while (0)

// cudaFuncAttributeMaxDynamicSharedMemorySize is sticky per (device, kernel): set it when a call needs more than any call
// before it on this device, not on every launch.  (A race between two host threads only repeats the call.)
#define CVM_SMEM_ATTR_ONCE(kernel, bytes)                                                                     \
    do {                                                                                                      \
        static size_t done_[64] = {0};                                                                        \
        int dev_ = 0;                                                                                         \
        CVM_CHECK_CUDA(cudaGetDevice(&dev_));                                                                 \
        if (dev_ < 0 || dev_ >= 64 || (size_t)(bytes) > done_[dev_]) {                                        \
            CVM_CHECK_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes))); \
            if (dev_ >= 0 && dev_ < 64) done_[dev_] = (size_t)(bytes);                                        \
        }                                                                                                     \
    } while (0)

static inline bool cvm_aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// ---- device: mbarrier + 1-D bulk async copy (TMA engine, SASS: UBLKCP / SYNCS) -----------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}

// make mbarrier.init visible to the async proxy before the first bulk copy signals it
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}

__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    const uint32_t addr = smem_u32(bar);
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(addr), "r"(parity)
            : "memory");
    } while (!ok);
}

__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// global -> shared bulk copy; bytes % 16 == 0, both addresses 16-byte aligned; completion counted on `bar`
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
        "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

// ---- device: reductions ------------------------------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// streaming stores / loads that do not pollute L1
__device__ __forceinline__ void st_cs_f4(float4* p, float4 v) {
    asm volatile("st.global.cs.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
