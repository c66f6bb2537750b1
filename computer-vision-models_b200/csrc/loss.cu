// Fused CenterNet / CenterTracker loss forward: ONE streaming pass over y_true and y_pred.
//
// Replaces the ~60 eager TF ops of CenternetLoss.call (reference models/centernet/loss.py:31-60,107-155) and
// CentertrackerLoss.track_offset_loss (models/centertracker/loss.py:16-28).  Bound: HBM.  Algorithmic bytes per pixel:
// 4*(y_true_stride + y_pred_stride).
//
// Structure (Blackwell): persistent CTAs; thread 0 streams spans of TP pixels of both tensors into a 3-stage shared
// memory ring with 1-D bulk async copies (cp.async.bulk -> UBLKCP, completion on an mbarrier), so no registers are
// tied up by loads in flight; all 256 threads consume a stage with a heatmap-element-flat mapping (bank conflicts
// <= 2-way for any channel stride).  Per-thread fp32 sums live for one span only and are folded into fp64 per-thread
// accumulators; warp shuffle -> per-block partials -> a fixed-order second kernel.  No float atomics: results are
// bit-reproducible run to run.
#include <cstdlib>

#include "common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kStages = 3;

struct LossParams {
    const float* yt;
    const float* yp;
    long long n_pixels;
    long long n_spans;
    int TP;            // pixels per span (multiple of 4)
    int st_t, st_p;    // per-pixel strides in floats
    int hm;
    int wch;           // weights channel inside a y_true pixel, -1 = unweighted (metric mode, loss.py:55-57)
    float inv_hm;
    float fa, fb;
    int a_is2, b_is4;
    int n_fields;
    int f_off[CVM_MAX_FIELDS], f_size[CVM_MAX_FIELDS], f_kind[CVM_MAX_FIELDS];
    int use_bulk;
    int n_stages;      // ring depth of the fast kernel
    double* block_partials;   // [gridDim.x][CVM_NPART]
    // the last block to finish sums the block partials in a fixed order (bit-reproducible) into `partials` and, when `fin_out`
    // is given (single-GPU callers: nothing to all-reduce), applies the finalise step right away: no extra launches
    unsigned int* ticket;     // zeroed by the host before the launch
    double* partials;         // [CVM_NPART]
    float* fin_out;           // [2 + CVM_MAX_FIELDS] or NULL
    int fin_post[CVM_MAX_FIELDS];
    float fin_weight[CVM_MAX_FIELDS];
};

__device__ __forceinline__ float pow_a(float x, const LossParams& p) { return p.a_is2 ? x * x : powf(x, p.fa); }
__device__ __forceinline__ float pow_b(float x, const LossParams& p) {
    if (p.b_is4) {
        float x2 = x * x;
        return x2 * x2;
    }
    return powf(x, p.fb);
}

// log(clip(1 - q, .01, .99)) of the negative term (loss.py:47).  z = clip(q) is exact, so for small z the series
// log1p(-z) = -z - z^2/2 - ... (6 terms, truncation < 1e-8 relative for z < 1/16) avoids both the rounding of 1 - q and
// the absolute-error floor of MUFU.LG2 near 1; elsewhere |log| >= 0.0645 and lg2.approx (abs err 2^-22) is < 3e-6 relative.
__device__ __forceinline__ float log_one_minus(float q) {
    const float z = fminf(fmaxf(q, 0.01f), 0.99f);
    float s = fmaf(z, 1.0f / 6.0f, 0.2f);
    s = fmaf(z, s, 0.25f);
    s = fmaf(z, s, 1.0f / 3.0f);
    s = fmaf(z, s, 0.5f);
    s = fmaf(z, s, 1.0f);
    const float small = -z * s;
    float lg;  // 1 - z is in [0.01, 0.99]: no denormal handling needed around MUFU.LG2
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(lg) : "f"(1.0f - z));
    const float big = lg * 0.69314718055994530942f;
    return z < 0.0625f ? small : big;
}

// one regression field at a peak pixel: sum_k |term| (loss.py:115-128), fp64 from fp32 inputs
__device__ double field_term(int kind, const float* t, const float* q, int size) {
    double s = 0.0;
    if (kind == CVM_KIND_CE) {
        // -sum_c t_c * log_softmax(q)_c, labels not renormalised (categorical_crossentropy from_logits, loss.py:122)
        double mx = -1e300;
        for (int k = 0; k < size; ++k) mx = fmax(mx, (double)q[k]);
        double se = 0.0;
        for (int k = 0; k < size; ++k) se += exp((double)q[k] - mx);
        const double lse = mx + log(se);
        for (int k = 0; k < size; ++k) s += (double)t[k] * (lse - (double)q[k]);
        return fabs(s);
    }
    for (int k = 0; k < size; ++k) {
        const double d = (double)t[k] - (double)q[k];
        if (kind == CVM_KIND_MSE)
            s += d * d;
        else if (kind == CVM_KIND_MAE)
            s += fabs(d);
        else
            s += fabs(d / fmax(fabs((double)t[k]), 1.0));
    }
    return s;
}

// loss.py:59 (tf.cond n>0), :130, :98 (orientation), :140-153 (weighting); one thread
__device__ void finalize_terms(const double* part, float* out, int n_fields, const int* f_post, const float* f_weight) {
    const double P = part[0], N = part[1], n = part[2], nobj = part[3];
    const double focal = n > 0.0 ? (P + N) / n : N;
    double total = focal;
    out[1] = (float)focal;
    for (int f = 0; f < n_fields; ++f) {
        double v = nobj > 0.0 ? part[4 + f] / nobj : part[4 + f];
        if (f_post[f] == CVM_POST_ORIENT) v = sqrt(1.0 - 0.99 * cos(2.0 * v)) + fabs(v * v * 0.05) - 0.0999;
        out[2 + f] = (float)v;
        total += v * (double)f_weight[f];
    }
    out[0] = (float)total;
}

// Called by every thread of a block after it has written its block partials: the last block of the grid to get here
// reduces all of them - 16 groups of blocks, then the groups, always in the same order, so the sums are bit-reproducible
// run to run whichever block comes last - and writes partials[CVM_NPART] (and the finalised terms).  NT = block size.
template <int NT>
__device__ __forceinline__ void last_block_reduce(const LossParams& p) {
    __shared__ double sh[16][CVM_NPART];
    __shared__ int s_last;
    __threadfence();            // this block's partials are visible device-wide before its ticket is
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(p.ticket, 1u) == gridDim.x - 1;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    const int n_blocks = (int)gridDim.x;
    const double* bp = p.block_partials;
    for (int t = threadIdx.x; t < 16 * CVM_NPART; t += NT) {
        const int k = t % CVM_NPART, g = t / CVM_NPART;
        double r = 0.0;
        // (L2 loads, eight in flight at a time; the additions stay in block order)
        for (int b0 = g; b0 < n_blocks; b0 += 16 * 8) {
            double v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int b = b0 + 16 * j;
                v[j] = b < n_blocks ? __ldcg(bp + (size_t)b * CVM_NPART + k) : 0.0;
            }
#pragma unroll
            for (int j = 0; j < 8; ++j)
                if (b0 + 16 * j < n_blocks) r += v[j];
        }
        sh[g][k] = r;
    }
    __syncthreads();
    if (threadIdx.x < CVM_NPART) {
        double t = 0.0;
        for (int gi = 0; gi < 16; ++gi) t += sh[gi][threadIdx.x];
        p.partials[threadIdx.x] = t;
        sh[0][threadIdx.x] = t;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        if (p.fin_out) finalize_terms(&sh[0][0], p.fin_out, p.n_fields, p.fin_post, p.fin_weight);
        *p.ticket = 0u;         // (the host zeroes it as well: a caller may hand over an uninitialised workspace)
    }
}

__global__ void __launch_bounds__(kThreads) loss_fwd_kernel(const LossParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ uint64_t full_bar[kStages];
    __shared__ double red[kThreads / 32][CVM_NPART];

    float* const ring = reinterpret_cast<float*>(smem_raw);
    const int tid = threadIdx.x;
    const int TP = p.TP;
    const size_t t_floats = (size_t)TP * p.st_t;
    const size_t stage_floats = (size_t)TP * (p.st_t + p.st_p);

    if (tid == 0) {
        for (int s = 0; s < kStages; ++s) mbar_init(&full_bar[s], 1);
        mbar_fence_init();
    }
    __syncthreads();

    // a span can use the bulk-copy engine when it is full (TP % 4 == 0 keeps 16-byte granularity) and pointers are aligned
    auto span_pixels = [&](long long span) -> int {
        const long long left = p.n_pixels - span * TP;
        return left < TP ? (int)left : TP;
    };
    auto span_is_bulk = [&](long long span) -> bool { return p.use_bulk && span_pixels(span) == TP; };
    auto load_span = [&](long long span, int s) {
        float* dst_t = ring + (size_t)s * stage_floats;
        float* dst_p = dst_t + t_floats;
        const float* src_t = p.yt + span * (long long)TP * p.st_t;
        const float* src_p = p.yp + span * (long long)TP * p.st_p;
        if (span_is_bulk(span)) {
            if (tid == 0) {
                const uint32_t bt = (uint32_t)(t_floats * 4), bp = (uint32_t)((size_t)TP * p.st_p * 4);
                mbar_arrive_expect_tx(&full_bar[s], bt + bp);
                bulk_g2s(dst_t, src_t, bt, &full_bar[s]);
                bulk_g2s(dst_p, src_p, bp, &full_bar[s]);
            }
        } else {  // ragged tail / unaligned tensors: cooperative copy, published by the __syncthreads in wait_span
            const int np = span_pixels(span);
            for (int i = tid; i < np * p.st_t; i += kThreads) dst_t[i] = src_t[i];
            for (int i = tid; i < np * p.st_p; i += kThreads) dst_p[i] = src_p[i];
        }
    };

    const long long gstride = gridDim.x;
    for (int s = 0; s < kStages; ++s) {
        const long long span = blockIdx.x + (long long)s * gstride;
        if (span < p.n_spans) load_span(span, s);
    }

    double accP = 0.0, accN = 0.0, accF[CVM_MAX_FIELDS];
#pragma unroll
    for (int f = 0; f < CVM_MAX_FIELDS; ++f) accF[f] = 0.0;
    unsigned int n_pos = 0, n_obj = 0;
    uint32_t phase_bits = 0;

    const int hm = p.hm, st_t = p.st_t, st_p = p.st_p, wch = p.wch;

    for (int it = 0;; ++it) {
        const long long span = blockIdx.x + (long long)it * gstride;
        if (span >= p.n_spans) break;
        const int s = it % kStages;
        if (span_is_bulk(span)) {
            mbar_wait(&full_bar[s], (phase_bits >> s) & 1u);
            phase_bits ^= (1u << s);
        } else {
            __syncthreads();
        }
        const float* __restrict__ T = ring + (size_t)s * stage_floats;
        const float* __restrict__ Q = T + t_floats;
        const int ne = span_pixels(span) * hm;

        float sP = 0.f, sN = 0.f;
#pragma unroll 4
        for (int j = tid; j < ne; j += kThreads) {
            int px, c;
            if (hm == 1) {
                px = j;
                c = 0;
            } else {
                px = __float2int_rz(((float)j + 0.5f) * p.inv_hm);
                c = j - px * hm;
            }
            const float y = T[px * st_t + c];
            const float q = Q[px * st_p + c];
            const float w = wch >= 0 ? T[px * st_t + wch] : 1.0f;
            if (y < 1.0f) {  // neg_mask (loss.py:36,43-48)
                sN += (-(pow_b(1.0f - y, p) * pow_a(q, p)) * log_one_minus(q)) * w;
            } else if (y == 1.0f) {  // pos_mask (loss.py:35,38-42)
                const float l = logf(fminf(fmaxf(q, 0.01f), 0.99f));
                sP += (-pow_a(1.0f - q, p) * l) * w;
                ++n_pos;
                // the lowest heatmap channel holding a 1.0 owns the pixel's regression terms (pos_mask reduce_max, loss.py:110-113)
                bool owner = true;
                for (int c2 = 0; c2 < c; ++c2) owner = owner && !(T[px * st_t + c2] == 1.0f);
                if (owner) {
                    ++n_obj;
#pragma unroll
                    for (int f = 0; f < CVM_MAX_FIELDS; ++f)
                        if (f < p.n_fields)
                            accF[f] += field_term(p.f_kind[f], T + px * st_t + p.f_off[f], Q + px * st_p + p.f_off[f],
                                                  p.f_size[f]);
                }
            }
        }
        accP += (double)sP;
        accN += (double)sN;

        __syncthreads();  // everyone is done with stage s
        const long long next = span + (long long)kStages * gstride;
        if (next < p.n_spans) load_span(next, s);
    }

    // block reduction: warp shuffles in fp64, then a fixed-order sum over the 8 warps
    double v[CVM_NPART];
    v[0] = accP;
    v[1] = accN;
    v[2] = (double)n_pos;
    v[3] = (double)n_obj;
#pragma unroll
    for (int f = 0; f < CVM_MAX_FIELDS; ++f) v[4 + f] = accF[f];
#pragma unroll
    for (int k = 4 + CVM_MAX_FIELDS; k < CVM_NPART; ++k) v[k] = 0.0;
    const int warp = tid >> 5, lane = tid & 31;
#pragma unroll
    for (int k = 0; k < CVM_NPART; ++k) {
        const double r = warp_sum(v[k]);
        if (lane == 0) red[warp][k] = r;
    }
    __syncthreads();
    if (tid < CVM_NPART) {
        double r = 0.0;
        for (int wi = 0; wi < kThreads / 32; ++wi) r += red[wi][tid];
        p.block_partials[(size_t)blockIdx.x * CVM_NPART + tid] = r;
    }
    last_block_reduce<kThreads>(p);
}

struct FinalizeParams {
    int n_fields;
    int f_post[CVM_MAX_FIELDS];
    float f_weight[CVM_MAX_FIELDS];
};

__global__ void loss_finalize_kernel(const double* __restrict__ part, float* __restrict__ out, const FinalizeParams fp) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    finalize_terms(part, out, fp.n_fields, fp.f_post, fp.f_weight);
}

int pick_span_pixels(int st_t, int st_p, size_t* smem_bytes) {
    // keep the 3-stage ring around <= 96 KB so two CTAs fit per SM
    int TP = 256;
    while (TP > 16 && (size_t)kStages * TP * (st_t + st_p) * 4 > 96 * 1024) TP >>= 1;
    *smem_bytes = (size_t)kStages * TP * (st_t + st_p) * 4;
    return TP;
}

int loss_grid(long long n_spans, size_t smem_bytes) {
    int per_sm = (int)((220 * 1024) / (smem_bytes + 2048));
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 4) per_sm = 4;
    long long g = (long long)cvm_num_sms() * per_sm;
    if (g > n_spans) g = n_spans;
    if (g < 1) g = 1;
    return (int)g;
}

// Rare work for a pixel that holds at least one heatmap value == 1.0: positive focal terms, counts and the regression
// fields.  Kept out of line, accumulating into a local-memory array, so that it costs the hot loop no registers.
// cold[0] = P, cold[1] = n_pos, cold[2] = n_obj, cold[3 + f] = field sums.
__device__ __noinline__ void peak_pixel(const float* T, const float* Q, int hm, float w, const LossParams& p, double* cold) {
    bool first = true;
    float sP = 0.f;
    for (int c = 0; c < hm; ++c) {
        if (T[c] == 1.0f) {                                                     // pos_mask, loss.py:35,38-42
            const float q = Q[c];
            sP += (-pow_a(1.0f - q, p) * logf(fminf(fmaxf(q, 0.01f), 0.99f))) * w;
            cold[1] += 1.0;
            if (first) {                                                        // pos_mask reduce_max, loss.py:110-113
                first = false;
                cold[2] += 1.0;
                for (int f = 0; f < p.n_fields; ++f)
                    cold[3 + f] += field_term(p.f_kind[f], T + p.f_off[f], Q + p.f_off[f], p.f_size[f]);
            }
        }
    }
    cold[0] += (double)sP;
}

// ---- packed fp32 pairs (sm_100: FFMA2 / FMUL2 / FADD2, PTX *.f32x2): two lanes per instruction, IEEE rounding per lane ----
typedef unsigned long long f2;
__device__ __forceinline__ f2 pk2(float a, float b) {
    f2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ void unpk2(f2 v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c) {
    f2 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
__device__ __forceinline__ f2 mul2(f2 a, f2 b) {
    f2 r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}

// ---- fast path: channel counts known at compile time, one pixel per thread ------------------------------------------
// All shared-memory offsets become immediates, the heatmap loop is fully unrolled (HM independent element chains per
// thread) and the rare "this pixel holds a peak" work is taken out of the hot loop.
// Fast forward kernel: 160 consumer threads (one pixel per thread and span) + the loader warp, and a ring as deep as lets two
// CTAs share an SM (5 stages for the CenterNet layouts).  Measured on B200 at the BASELINE shape: 256 threads x 3 stages
// 0.260 ms, 192 x 4 0.248 ms, 160 x 5 0.245 ms: bytes in flight matter more than consumer warps.
constexpr int kFastThreads = 160;
constexpr int kFastStages = 5;    // most stages the barrier arrays hold; LossParams::n_stages <= this are used

template <int HM, int ST_T, int ST_P>
__global__ void __launch_bounds__(kFastThreads + 32) loss_fwd_fast_kernel(const LossParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ uint64_t full_bar[kFastStages];    // loader -> consumers: span arrived
    __shared__ uint64_t empty_bar[kFastStages];   // consumers -> loader: all consumer warps are done with the stage
    __shared__ double red[kFastThreads / 32][CVM_NPART];

    float* const ring = reinterpret_cast<float*>(smem_raw);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n_stages = p.n_stages;
    constexpr int TP = kFastThreads;  // one pixel per consumer thread per span
    constexpr int kConsumerWarps = kFastThreads / 32;
    constexpr size_t t_floats = (size_t)TP * ST_T;
    constexpr size_t stage_floats = (size_t)TP * (ST_T + ST_P);

    if (tid == 0) {
        for (int s = 0; s < n_stages; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], kConsumerWarps);
        }
        mbar_fence_init();
    }
    __syncthreads();   // the only CTA-wide barrier before the final reduction: loader and consumers run on their own

    const long long n_full = p.use_bulk ? p.n_pixels / TP : 0;  // spans below this index are full and bulk-loadable
    const long long gstride = gridDim.x;
    double v[CVM_NPART];
#pragma unroll
    for (int k = 0; k < CVM_NPART; ++k) v[k] = 0.0;

    if (warp == kConsumerWarps) {
        // ---- loader warp: span `it` of this CTA goes to stage it % n_stages once every consumer warp has released it ----
        int s = 0;
        uint32_t e_parity = 0;
        for (int it = 0;; ++it) {
            // spans are walked from the END of the tensors: y_true has just been written by the render in the usual pipeline,
            // and the last ~100 MB of it are still in L2 (126 MB) - read those first, before the stream evicts them
            const long long span = p.n_spans - 1 - (blockIdx.x + (long long)it * gstride);
            if (span < 0) break;
            if (it >= n_stages) mbar_wait(&empty_bar[s], e_parity);
            float* dst_t = ring + (size_t)s * stage_floats;
            float* dst_p = dst_t + t_floats;
            const float* src_t = p.yt + span * (long long)(TP * ST_T);
            const float* src_p = p.yp + span * (long long)(TP * ST_P);
            if (span < n_full) {
                if (lane == 0) {
                    constexpr uint32_t bt = (uint32_t)(t_floats * 4), bq = (uint32_t)((size_t)TP * ST_P * 4);
                    mbar_arrive_expect_tx(&full_bar[s], bt + bq);
                    bulk_g2s(dst_t, src_t, bt, &full_bar[s]);
                    bulk_g2s(dst_p, src_p, bq, &full_bar[s]);
                }
            } else {   // ragged last span / unaligned tensors: plain loads by the loader warp
                const long long left = p.n_pixels - span * TP;
                const int np = left < TP ? (int)left : TP;
                for (int i = lane; i < np * ST_T; i += 32) dst_t[i] = src_t[i];
                for (int i = lane; i < np * ST_P; i += 32) dst_p[i] = src_p[i];
                __syncwarp();
                if (lane == 0) mbar_arrive(&full_bar[s]);   // release: the stores above are visible to the waiters
            }
            if (++s == n_stages) {
                s = 0;
                if (it >= n_stages) e_parity ^= 1u;
            }
        }
    } else {
        // ---- consumer warps ----
        double accN = 0.0;
        double cold[3 + CVM_MAX_FIELDS];
#pragma unroll
        for (int f = 0; f < 3 + CVM_MAX_FIELDS; ++f) cold[f] = 0.0;
        const int wch = p.wch;
        float sN = 0.f;
        int folded = 0, s = 0;
        uint32_t f_parity = 0;
        for (int it = 0;; ++it) {
            // spans are walked from the END of the tensors: y_true has just been written by the render in the usual pipeline,
            // and the last ~100 MB of it are still in L2 (126 MB) - read those first, before the stream evicts them
            const long long span = p.n_spans - 1 - (blockIdx.x + (long long)it * gstride);
            if (span < 0) break;
            const long long left = p.n_pixels - span * TP;
            const int np = left < TP ? (int)left : TP;
            mbar_wait(&full_bar[s], f_parity);
            const float* __restrict__ T = ring + (size_t)s * stage_floats + tid * ST_T;
            const float* __restrict__ Q = ring + (size_t)s * stage_floats + t_floats + tid * ST_P;
            if (tid < np) {
                const float w = wch >= 0 ? T[wch] : 1.0f;
                float ymax = 0.f;
                // Two heatmap channels per instruction: sm_100's packed fp32 pipe (FFMA2/FMUL2, PTX *.f32x2) runs the
                // multiplies and the log1p series of a channel pair in one issue slot each; per lane the operations and
                // their rounding are those of the scalar code (the kernel is issue bound, not HBM bound).  acc2 holds
                // +neg_loss of the even / odd channels.  For y <= 1 the neg_mask (loss.py:36) is implied by (1-y)^4 = 0
                // at y == 1; a pixel with some y > 1 (never produced by the render) is redone exactly below.
                f2 acc2 = pk2(0.f, 0.f);
#pragma unroll
                for (int c = 0; c + 1 < HM; c += 2) {
                    const float y0 = T[c], y1 = T[c + 1];
                    float q0, q1;
                    f2 q;
                    if (ST_P % 2 == 0) {
                        q = *reinterpret_cast<const f2*>(Q + c);   // 8-byte aligned: even pixel stride, even channel
                        unpk2(q, q0, q1);
                    } else {
                        q0 = Q[c];
                        q1 = Q[c + 1];
                        q = pk2(q0, q1);
                    }
                    const f2 t = fma2(pk2(y0, y1), pk2(-1.f, -1.f), pk2(1.f, 1.f));               // 1 - y
                    const f2 t2 = mul2(t, t);                                                      // alpha = 2, beta = 4 only
                    const f2 tq = mul2(mul2(t2, t2), mul2(q, q));
                    // -log(clip(1 - q, .01, .99)) (loss.py:47), see log_one_minus(): series below 1/16, MUFU.LG2 above
                    const float z0 = fminf(fmaxf(q0, 0.01f), 0.99f), z1 = fminf(fmaxf(q1, 0.01f), 0.99f);
                    const f2 z = pk2(z0, z1);
                    f2 sr = fma2(z, pk2(1.0f / 6.0f, 1.0f / 6.0f), pk2(0.2f, 0.2f));
                    sr = fma2(z, sr, pk2(0.25f, 0.25f));
                    sr = fma2(z, sr, pk2(1.0f / 3.0f, 1.0f / 3.0f));
                    sr = fma2(z, sr, pk2(0.5f, 0.5f));
                    sr = fma2(z, sr, pk2(1.0f, 1.0f));
                    float s0, s1;
                    unpk2(mul2(z, sr), s0, s1);                                                    // -log1p(-z) for small z
                    float lg0, lg1;
                    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(lg0) : "f"(1.0f - z0));
                    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(lg1) : "f"(1.0f - z1));
                    const float l0 = z0 < 0.0625f ? s0 : lg0 * -0.69314718055994530942f;
                    const float l1 = z1 < 0.0625f ? s1 : lg1 * -0.69314718055994530942f;
                    acc2 = fma2(tq, pk2(l0, l1), acc2);                                            // +neg_loss, loss.py:43-48
                    ymax = fmaxf(ymax, fmaxf(y0, y1));
                }
                float a0, a1;
                unpk2(acc2, a0, a1);
                float acc = a0 + a1;
                if (HM & 1) {   // odd channel count: the last channel on the scalar path
                    const float y = T[HM - 1], q = Q[HM - 1];
                    const float t = 1.0f - y, t2 = t * t;
                    acc -= ((t2 * t2) * (q * q)) * log_one_minus(q);
                    ymax = fmaxf(ymax, y);
                }
                if (ymax > 1.0f) {   // out-of-range targets: the masked sum, channel by channel
                    acc = 0.f;
                    for (int c = 0; c < HM; ++c) {
                        const float y = T[c], q = Q[c];
                        const float t = 1.0f - y, t2 = t * t;
                        if (y < 1.0f) acc -= ((t2 * t2) * (q * q)) * log_one_minus(q);
                    }
                }
                sN = fmaf(acc, w, sN);
                if (ymax >= 1.0f) peak_pixel(T, Q, HM, w, p, cold);  // rare: a few dozen pixels per image (re-tests == 1.0)
            }
            if (++folded == 8) {  // short fp32 chains (<= 8*HM addends), everything above in fp64
                accN += (double)sN;
                sN = 0.f;
                folded = 0;
            }
            __syncwarp();   // every lane is done reading the stage
            if (lane == 0) mbar_arrive(&empty_bar[s]);
            if (++s == n_stages) {
                s = 0;
                f_parity ^= 1u;
            }
        }
        accN += (double)sN;
        v[0] = cold[0];
        v[1] = accN;
        v[2] = cold[1];
        v[3] = cold[2];
#pragma unroll
        for (int f = 0; f < CVM_MAX_FIELDS; ++f) v[4 + f] = cold[3 + f];
#pragma unroll
        for (int k = 0; k < CVM_NPART; ++k) {
            const double r = warp_sum(v[k]);
            if (lane == 0) red[warp][k] = r;
        }
    }
    __syncthreads();
    if (tid < CVM_NPART) {
        double r = 0.0;
        for (int wi = 0; wi < kConsumerWarps; ++wi) r += red[wi][tid];
        p.block_partials[(size_t)blockIdx.x * CVM_NPART + tid] = r;
    }
    last_block_reduce<kFastThreads + 32>(p);
}

template <int HM, int ST_T, int ST_P>
int launch_loss_fast(LossParams& p, cudaStream_t st, int* grid_out) {
    p.TP = kFastThreads;
    p.n_spans = (p.n_pixels + p.TP - 1) / p.TP;
    const size_t stage_bytes = (size_t)kFastThreads * (ST_T + ST_P) * 4;
    int n_stages = (int)((size_t)(110 * 1024) / stage_bytes);   // two CTAs per SM
    if (n_stages > kFastStages) n_stages = kFastStages;
    if (n_stages < 2) n_stages = 2;
    p.n_stages = n_stages;
    const size_t smem = (size_t)n_stages * stage_bytes;
    const int grid = loss_grid(p.n_spans, smem);
    CVM_SMEM_ATTR_ONCE((loss_fwd_fast_kernel<HM, ST_T, ST_P>), smem);
    loss_fwd_fast_kernel<HM, ST_T, ST_P><<<grid, kFastThreads + 32, smem, st>>>(p);
    CVM_CHECK_LAUNCH("loss_fwd_fast_kernel");
    *grid_out = grid;
    return CVM_OK;
}

int check_layout_for_loss(const cvm_layout* L, int st_t, int st_p) {
    CVM_CHECK_ARG(L != nullptr, "layout is NULL");
    CVM_CHECK_ARG(L->hm >= 1 && L->hm <= 64, "hm=%d out of range [1,64]", L->hm);
    CVM_CHECK_ARG(L->Cp >= L->hm && L->Ct == L->Cp + 1, "need Cp >= hm and Ct == Cp+1 (Cp=%d Ct=%d hm=%d)", L->Cp, L->Ct, L->hm);
    CVM_CHECK_ARG(st_t >= L->Ct && st_p >= L->Cp, "strides (%d,%d) smaller than channel counts (%d,%d)", st_t, st_p, L->Ct, L->Cp);
    CVM_CHECK_ARG(st_t <= 256 && st_p <= 256, "pixel strides above 256 floats are not supported");
    CVM_CHECK_ARG(L->n_fields >= 0 && L->n_fields <= CVM_MAX_FIELDS, "n_fields=%d", L->n_fields);
    for (int f = 0; f < L->n_fields; ++f) {
        CVM_CHECK_ARG(L->field_off[f] >= L->hm && L->field_off[f] + L->field_size[f] <= L->Cp && L->field_size[f] >= 1,
                      "field %d [%d,%d) outside [hm,Cp)", f, L->field_off[f], L->field_off[f] + L->field_size[f]);
        CVM_CHECK_ARG(L->field_kind[f] >= 0 && L->field_kind[f] <= 3, "field %d: unknown kind %d", f, L->field_kind[f]);
    }
    return CVM_OK;
}

}  // namespace

extern "C" size_t cvm_loss_workspace_bytes(const cvm_layout* L, long long n_pixels) {
    (void)L;
    (void)n_pixels;
    // per-block partials for the largest grid we ever launch (4 CTAs/SM) + the ticket of the last-block reduction
    return (size_t)cvm_num_sms() * 4 * CVM_NPART * sizeof(double) + 16;
}

namespace {
int loss_fwd_impl(const cvm_layout* L, const float* y_true, int y_true_stride, const float* y_pred, int y_pred_stride,
                  long long n_pixels, int use_weights, double* partials, float* fin_out, void* ws, size_t ws_bytes, void* stream) {
    int rc = check_layout_for_loss(L, y_true_stride, y_pred_stride);
    if (rc != CVM_OK) return rc;
    CVM_CHECK_ARG(y_true && y_pred && partials && ws, "NULL pointer argument");
    CVM_CHECK_ARG(n_pixels >= 0, "n_pixels < 0");
    if (ws_bytes < cvm_loss_workspace_bytes(L, n_pixels)) {
        cvm_set_error("workspace too small: %zu < %zu", ws_bytes, cvm_loss_workspace_bytes(L, n_pixels));
        return CVM_ERR_WS;
    }
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);

    LossParams p;
    memset(&p, 0, sizeof(p));
    size_t smem = 0;
    p.TP = pick_span_pixels(y_true_stride, y_pred_stride, &smem);
    p.yt = y_true;
    p.yp = y_pred;
    p.n_pixels = n_pixels;
    p.n_spans = (n_pixels + p.TP - 1) / p.TP;
    p.st_t = y_true_stride;
    p.st_p = y_pred_stride;
    p.hm = L->hm;
    p.wch = use_weights ? L->Ct - 1 : -1;
    p.inv_hm = 1.0f / (float)L->hm;
    p.fa = L->focal_a;
    p.fb = L->focal_b;
    p.a_is2 = (L->focal_a == 2.0f);
    p.b_is4 = (L->focal_b == 4.0f);
    p.n_fields = L->n_fields;
    for (int f = 0; f < L->n_fields; ++f) {
        p.f_off[f] = L->field_off[f];
        p.f_size[f] = L->field_size[f];
        p.f_kind[f] = L->field_kind[f];
    }
    p.use_bulk = cvm_aligned16(y_true) && cvm_aligned16(y_pred);
    p.block_partials = static_cast<double*>(ws);
    p.ticket = reinterpret_cast<unsigned int*>(static_cast<unsigned char*>(ws) + (size_t)cvm_num_sms() * 4 * CVM_NPART * sizeof(double));
    p.partials = partials;
    p.fin_out = fin_out;
    for (int f = 0; f < L->n_fields; ++f) {
        p.fin_post[f] = L->field_post[f];
        p.fin_weight[f] = L->field_weight[f];
    }
    CVM_CHECK_CUDA(cudaMemsetAsync(p.ticket, 0, 4, st));   // (a memset node, not a launch; the workspace may be uninitialised)

    int grid = 0;
    // compile-time layouts: (hm, y_true stride, y_pred stride) of BASELINE.json's configs and the reference defaults
#define CVM_LOSS_FAST(HM_, ST_T_, ST_P_)                                              \
    if (grid == 0 && p.a_is2 && p.b_is4 && L->hm == HM_ && y_true_stride == ST_T_ && y_pred_stride == ST_P_) { \
        rc = launch_loss_fast<HM_, ST_T_, ST_P_>(p, st, &grid);                        \
        if (rc != CVM_OK) return rc;                                                  \
    }
    CVM_LOSS_FAST(10, 15, 14)   // configs 1-3: 10-class heatmaps
    CVM_LOSS_FAST(10, 17, 16)   // config 4: + track_offset
    CVM_LOSS_FAST(10, 22, 20)   // config 5: CenterNet slice inside the multitask tensors
    CVM_LOSS_FAST(1, 16, 15)    // reference-exact layout, 10 classes
    CVM_LOSS_FAST(1, 12, 11)    // reference defaults (6 classes)
    CVM_LOSS_FAST(1, 14, 13)    // reference CenterTracker defaults
#undef CVM_LOSS_FAST
    if (grid == 0) {            // any other layout: runtime-parameterised kernel
        grid = loss_grid(p.n_spans, smem);
        CVM_SMEM_ATTR_ONCE((loss_fwd_kernel), smem);
        loss_fwd_kernel<<<grid, kThreads, smem, st>>>(p);
        CVM_CHECK_LAUNCH("loss_fwd_kernel");
    }
    return CVM_OK;
}
}  // namespace

extern "C" int cvm_loss_fwd(const cvm_layout* L, const float* y_true, int y_true_stride, const float* y_pred,
                            int y_pred_stride, long long n_pixels, int use_weights, double* partials, void* ws,
                            size_t ws_bytes, void* stream) {
    return loss_fwd_impl(L, y_true, y_true_stride, y_pred, y_pred_stride, n_pixels, use_weights, partials, nullptr, ws, ws_bytes, stream);
}

extern "C" int cvm_loss_fwd_total(const cvm_layout* L, const float* y_true, int y_true_stride, const float* y_pred,
                                  int y_pred_stride, long long n_pixels, int use_weights, double* partials, float* out, void* ws,
                                  size_t ws_bytes, void* stream) {
    CVM_CHECK_ARG(out != nullptr, "out is NULL");
    return loss_fwd_impl(L, y_true, y_true_stride, y_pred, y_pred_stride, n_pixels, use_weights, partials, out, ws, ws_bytes, stream);
}

namespace {
// partials of n_ranks shards, summed in rank order (bit-reproducible whatever algorithm gathered them), then finalised
__global__ void loss_finalize_gathered_kernel(const double* __restrict__ gathered, int n_ranks, double* __restrict__ partials,
                                              float* __restrict__ out, const FinalizeParams fp) {
    __shared__ double sum[CVM_NPART];
    if (threadIdx.x < CVM_NPART) {
        double t = 0.0;
        for (int r = 0; r < n_ranks; ++r) t += gathered[(size_t)r * CVM_NPART + threadIdx.x];
        sum[threadIdx.x] = t;
        if (partials) partials[threadIdx.x] = t;
    }
    __syncthreads();
    if (threadIdx.x == 0 && out) finalize_terms(sum, out, fp.n_fields, fp.f_post, fp.f_weight);
}
}  // namespace

extern "C" int cvm_loss_finalize_gathered(const cvm_layout* L, const double* gathered, int n_ranks, double* partials, float* out,
                                          void* stream) {
    CVM_CHECK_ARG(L && gathered && (partials || out), "NULL pointer argument");
    CVM_CHECK_ARG(n_ranks >= 1, "n_ranks < 1");
    CVM_CHECK_ARG(L->n_fields >= 0 && L->n_fields <= CVM_MAX_FIELDS, "n_fields=%d", L->n_fields);
    FinalizeParams fp;
    memset(&fp, 0, sizeof(fp));
    fp.n_fields = L->n_fields;
    for (int f = 0; f < L->n_fields; ++f) {
        fp.f_post[f] = L->field_post[f];
        fp.f_weight[f] = L->field_weight[f];
    }
    loss_finalize_gathered_kernel<<<1, 32, 0, reinterpret_cast<cudaStream_t>(stream)>>>(gathered, n_ranks, partials, out, fp);
    CVM_CHECK_LAUNCH("loss_finalize_gathered_kernel");
    return CVM_OK;
}

extern "C" int cvm_loss_finalize(const cvm_layout* L, const double* partials, float* out, void* stream) {
    CVM_CHECK_ARG(L && partials && out, "NULL pointer argument");
    CVM_CHECK_ARG(L->n_fields >= 0 && L->n_fields <= CVM_MAX_FIELDS, "n_fields=%d", L->n_fields);
    FinalizeParams fp;
    memset(&fp, 0, sizeof(fp));
    fp.n_fields = L->n_fields;
    for (int f = 0; f < L->n_fields; ++f) {
        fp.f_post[f] = L->field_post[f];
        fp.f_weight[f] = L->field_weight[f];
    }
    loss_finalize_kernel<<<1, 32, 0, reinterpret_cast<cudaStream_t>(stream)>>>(partials, out, fp);
    CVM_CHECK_LAUNCH("loss_finalize_kernel");
    return CVM_OK;
}

// =====================================================================================================================
// Backward: grad_pred[n_pixels, Cp] = upstream * d(total)/d(y_pred)   (TF autograd through loss.py in the reference)
// Same streaming structure as the forward; the gradient span is built in shared memory (zeros + heatmap terms + the few
// regression terms at peak pixels) and leaves with one bulk async store (shared -> global) per span.
// =====================================================================================================================
namespace {

struct BwdParams {
    LossParams fwd;                 // tensors, strides, fields (block_partials unused)
    const double* partials;         // globally reduced [P, N, n_pos, n_obj, field sums]
    const float* upstream;          // device scalar or nullptr
    float* grad;                    // [n_pixels, Cp]
    int Cp;
    int f_post[CVM_MAX_FIELDS];
    float f_weight[CVM_MAX_FIELDS];
    int grad_bulk;
};

__device__ __forceinline__ float dpow(float x, float a, int is2) {  // d/dx x^a
    return is2 ? 2.0f * x : a * powf(x, a - 1.0f);
}

__device__ __forceinline__ void bulk_s2g(void* dst_gmem, const void* src_smem, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem), "r"(smem_u32(src_smem)),
                 "r"(bytes)
                 : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}

__global__ void __launch_bounds__(kThreads) loss_bwd_kernel(const BwdParams bp) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ uint64_t full_bar[kStages];
    __shared__ float s_fscale[CVM_MAX_FIELDS];
    __shared__ float s_focal_scale;

    const LossParams& p = bp.fwd;
    float* const ring = reinterpret_cast<float*>(smem_raw);
    const int tid = threadIdx.x;
    const int TP = p.TP, Cp = bp.Cp;
    const size_t t_floats = (size_t)TP * p.st_t;
    const size_t stage_floats = (size_t)TP * (p.st_t + p.st_p);
    float* const gbuf = ring + (size_t)kStages * stage_floats;  // 2 x [TP*Cp]
    const size_t g_floats = (size_t)TP * Cp;

    if (tid == 0) {
        for (int s = 0; s < kStages; ++s) mbar_init(&full_bar[s], 1);
        mbar_fence_init();
        const double up = bp.upstream ? (double)*bp.upstream : 1.0;
        const double n = bp.partials[2], nobj = bp.partials[3];
        s_focal_scale = (float)(n > 0.0 ? up / n : up);                                  // loss.py:59
        for (int f = 0; f < p.n_fields; ++f) {
            double sc = up * (double)bp.f_weight[f] * (nobj > 0.0 ? 1.0 / nobj : 1.0);    // loss.py:130,140-153
            if (bp.f_post[f] == CVM_POST_ORIENT) {                                        // loss.py:98
                const double v = nobj > 0.0 ? bp.partials[4 + f] / nobj : bp.partials[4 + f];
                sc *= (0.99 * sin(2.0 * v)) / sqrt(1.0 - 0.99 * cos(2.0 * v)) + 0.1 * v;
            }
            s_fscale[f] = (float)sc;
        }
    }
    __syncthreads();

    auto span_pixels = [&](long long span) -> int {
        const long long left = p.n_pixels - span * TP;
        return left < TP ? (int)left : TP;
    };
    auto span_is_bulk = [&](long long span) -> bool { return p.use_bulk && span_pixels(span) == TP; };
    auto load_span = [&](long long span, int s) {
        float* dst_t = ring + (size_t)s * stage_floats;
        float* dst_p = dst_t + t_floats;
        const float* src_t = p.yt + span * (long long)TP * p.st_t;
        const float* src_p = p.yp + span * (long long)TP * p.st_p;
        if (span_is_bulk(span)) {
            if (tid == 0) {
                const uint32_t bt = (uint32_t)(t_floats * 4), bq = (uint32_t)((size_t)TP * p.st_p * 4);
                mbar_arrive_expect_tx(&full_bar[s], bt + bq);
                bulk_g2s(dst_t, src_t, bt, &full_bar[s]);
                bulk_g2s(dst_p, src_p, bq, &full_bar[s]);
            }
        } else {
            const int np = span_pixels(span);
            for (int i = tid; i < np * p.st_t; i += kThreads) dst_t[i] = src_t[i];
            for (int i = tid; i < np * p.st_p; i += kThreads) dst_p[i] = src_p[i];
        }
    };

    const long long gstride = gridDim.x;
    for (int s = 0; s < kStages; ++s) {
        const long long span = blockIdx.x + (long long)s * gstride;
        if (span < p.n_spans) load_span(span, s);
    }

    const int hm = p.hm, st_t = p.st_t, st_p = p.st_p, wch = p.wch;
    const float fscale = s_focal_scale;
    uint32_t phase_bits = 0;

    for (int it = 0;; ++it) {
        const long long span = blockIdx.x + (long long)it * gstride;
        if (span >= p.n_spans) break;
        const int s = it % kStages;
        float* const G = gbuf + (size_t)(it & 1) * g_floats;
        // the bulk store issued two iterations ago must have finished READING this gradient buffer
        if (tid == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        if (span_is_bulk(span)) {
            mbar_wait(&full_bar[s], (phase_bits >> s) & 1u);
            phase_bits ^= (1u << s);
        }
        __syncthreads();
        const int np = span_pixels(span);
        for (int i = tid; i < np * Cp; i += kThreads) G[i] = 0.f;
        __syncthreads();

        const float* __restrict__ T = ring + (size_t)s * stage_floats;
        const float* __restrict__ Q = T + t_floats;
        const int ne = np * hm;
        for (int j = tid; j < ne; j += kThreads) {
            int px, c;
            if (hm == 1) {
                px = j;
                c = 0;
            } else {
                px = __float2int_rz(((float)j + 0.5f) * p.inv_hm);
                c = j - px * hm;
            }
            const float y = T[px * st_t + c];
            const float q = Q[px * st_p + c];
            const float w = wch >= 0 ? T[px * st_t + wch] : 1.0f;
            float g = 0.f;
            if (y < 1.0f) {
                const float u = 1.0f - q;
                const float l = logf(fminf(fmaxf(u, 0.01f), 0.99f));
                const float dl = (u >= 0.01f && u <= 0.99f) ? -1.0f / u : 0.0f;   // clip_by_value passes grad inside [min,max]
                g = -pow_b(1.0f - y, p) * (dpow(q, p.fa, p.a_is2) * l + pow_a(q, p) * dl);
            } else if (y == 1.0f) {
                const float u = 1.0f - q;
                const float l = logf(fminf(fmaxf(q, 0.01f), 0.99f));
                const float dl = (q >= 0.01f && q <= 0.99f) ? 1.0f / q : 0.0f;
                g = dpow(u, p.fa, p.a_is2) * l - pow_a(u, p) * dl;
                bool owner = true;
                for (int c2 = 0; c2 < c; ++c2) owner = owner && !(T[px * st_t + c2] == 1.0f);
                if (owner) {
                    for (int f = 0; f < p.n_fields; ++f) {
                        const float* t = T + px * st_t + p.f_off[f];
                        const float* qq = Q + px * st_p + p.f_off[f];
                        float* gg = G + px * Cp + p.f_off[f];
                        const int size = p.f_size[f], kind = p.f_kind[f];
                        const float sc = s_fscale[f];
                        if (kind == CVM_KIND_CE) {
                            float mx = qq[0];
                            for (int k = 1; k < size; ++k) mx = fmaxf(mx, qq[k]);
                            float se = 0.f, st = 0.f;
                            for (int k = 0; k < size; ++k) {
                                se += expf(qq[k] - mx);
                                st += t[k];
                            }
                            const float lse = mx + logf(se);
                            float ce = 0.f;
                            for (int k = 0; k < size; ++k) ce += t[k] * (lse - qq[k]);
                            const float sg = ce > 0.f ? 1.f : (ce < 0.f ? -1.f : 0.f);
                            for (int k = 0; k < size; ++k) gg[k] += sc * sg * (st * expf(qq[k] - lse) - t[k]);
                        } else {
                            for (int k = 0; k < size; ++k) {
                                const float d = t[k] - qq[k];
                                float dv;
                                if (kind == CVM_KIND_MSE)
                                    dv = -2.0f * d;
                                else if (kind == CVM_KIND_MAE)
                                    dv = d > 0.f ? -1.f : (d < 0.f ? 1.f : 0.f);
                                else {
                                    const float den = fmaxf(fabsf(t[k]), 1.0f);
                                    dv = (d > 0.f ? -1.f : (d < 0.f ? 1.f : 0.f)) / den;
                                }
                                gg[k] += sc * dv;
                            }
                        }
                    }
                }
            }
            G[px * Cp + c] = g * w * fscale;
        }
        __syncthreads();
        float* dst = bp.grad + span * (long long)TP * Cp;
        if (bp.grad_bulk && np == TP) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> async-proxy read
            __syncthreads();
            if (tid == 0) bulk_s2g(dst, G, (uint32_t)(g_floats * 4));
        } else {
            for (int i = tid; i < np * Cp; i += kThreads) dst[i] = G[i];
            __syncthreads();
        }
        const long long next = span + (long long)kStages * gstride;
        if (next < p.n_spans) load_span(next, s);
    }
    if (tid == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // smem must outlive the last store's read
}

// ---- fast backward: channel counts known at compile time, one pixel per consumer thread ------------------------------
// Same pipeline as the fast forward kernel plus a third shared-memory buffer per stage for the gradients: the loader warp
// fetches y_true / y_pred spans with bulk async copies, the consumer warps write d(total)/d(y_pred) of their pixels into
// the stage's gradient buffer, and the loader warp streams that buffer out with one bulk store per span.  No CTA-wide
// barrier; alpha = 2, beta = 4 only (anything else takes the generic kernel).
constexpr int kBwdThreads = 160;   // consumer threads (one pixel per thread and span)
constexpr int kBwdStages = 5;      // most stages the barrier arrays hold

// One pixel that holds a target == 1 (or > 1): positive focal terms and the regression fields, element by element like
// the generic kernel.  Rare: a few dozen pixels per image.
__device__ __noinline__ void bwd_peak_pixel(const LossParams& p, const BwdParams& bp, const float* T, const float* Q, float* G,
                                            float w, float fscale, const float* fsc) {
    const int hm = p.hm;
    for (int c = 0; c < bp.Cp; ++c) G[c] = 0.f;
    for (int c = 0; c < hm; ++c) {
        const float y = T[c], q = Q[c];
        float g = 0.f;
        if (y < 1.0f) {
            const float u = 1.0f - q;
            const float l = logf(fminf(fmaxf(u, 0.01f), 0.99f));
            const float dl = (u >= 0.01f && u <= 0.99f) ? -1.0f / u : 0.0f;
            g = -pow_b(1.0f - y, p) * (dpow(q, p.fa, p.a_is2) * l + pow_a(q, p) * dl);
        } else if (y == 1.0f) {
            const float u = 1.0f - q;
            const float l = logf(fminf(fmaxf(q, 0.01f), 0.99f));
            const float dl = (q >= 0.01f && q <= 0.99f) ? 1.0f / q : 0.0f;
            g = dpow(u, p.fa, p.a_is2) * l - pow_a(u, p) * dl;
            bool owner = true;
            for (int c2 = 0; c2 < c; ++c2) owner = owner && !(T[c2] == 1.0f);
            if (owner) {
                for (int f = 0; f < p.n_fields; ++f) {
                    const float* t = T + p.f_off[f];
                    const float* qq = Q + p.f_off[f];
                    float* gg = G + p.f_off[f];
                    const int size = p.f_size[f], kind = p.f_kind[f];
                    const float sc = fsc[f];
                    if (kind == CVM_KIND_CE) {
                        float mx = qq[0];
                        for (int k = 1; k < size; ++k) mx = fmaxf(mx, qq[k]);
                        float se = 0.f, st = 0.f;
                        for (int k = 0; k < size; ++k) {
                            se += expf(qq[k] - mx);
                            st += t[k];
                        }
                        const float lse = mx + logf(se);
                        float ce = 0.f;
                        for (int k = 0; k < size; ++k) ce += t[k] * (lse - qq[k]);
                        const float sg = ce > 0.f ? 1.f : (ce < 0.f ? -1.f : 0.f);
                        for (int k = 0; k < size; ++k) gg[k] += sc * sg * (st * expf(qq[k] - lse) - t[k]);
                    } else {
                        for (int k = 0; k < size; ++k) {
                            const float d = t[k] - qq[k];
                            float dv;
                            if (kind == CVM_KIND_MSE)
                                dv = -2.0f * d;
                            else if (kind == CVM_KIND_MAE)
                                dv = d > 0.f ? -1.f : (d < 0.f ? 1.f : 0.f);
                            else
                                dv = (d > 0.f ? -1.f : (d < 0.f ? 1.f : 0.f)) / fmaxf(fabsf(t[k]), 1.0f);
                            gg[k] += sc * dv;
                        }
                    }
                }
            }
        }
        G[c] = g * w * fscale;
    }
}

template <int HM, int ST_T, int ST_P, int CP>
__global__ void __launch_bounds__(kBwdThreads + 32) loss_bwd_fast_kernel(const BwdParams bp) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ uint64_t full_bar[kBwdStages];   // loader -> consumers: span arrived
    __shared__ uint64_t done_bar[kBwdStages];   // consumers -> loader: the stage's gradients are complete (and its inputs read)
    __shared__ float s_fscale[CVM_MAX_FIELDS];
    __shared__ float s_focal_scale;

    const LossParams& p = bp.fwd;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n_stages = p.n_stages;
    constexpr int TP = kBwdThreads;
    constexpr int kConsumerWarps = kBwdThreads / 32;
    constexpr size_t t_floats = (size_t)TP * ST_T, q_floats = (size_t)TP * ST_P, g_floats = (size_t)TP * CP;
    constexpr size_t stage_floats = t_floats + q_floats + g_floats;
    float* const ring = reinterpret_cast<float*>(smem_raw);

    if (tid == 0) {
        for (int s = 0; s < n_stages; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&done_bar[s], kConsumerWarps);
        }
        mbar_fence_init();
        const double up = bp.upstream ? (double)*bp.upstream : 1.0;
        const double n = bp.partials[2], nobj = bp.partials[3];
        s_focal_scale = (float)(n > 0.0 ? up / n : up);                                  // loss.py:59
        for (int f = 0; f < p.n_fields; ++f) {
            double sc = up * (double)bp.f_weight[f] * (nobj > 0.0 ? 1.0 / nobj : 1.0);    // loss.py:130,140-153
            if (bp.f_post[f] == CVM_POST_ORIENT) {                                        // loss.py:98
                const double v = nobj > 0.0 ? bp.partials[4 + f] / nobj : bp.partials[4 + f];
                sc *= (0.99 * sin(2.0 * v)) / sqrt(1.0 - 0.99 * cos(2.0 * v)) + 0.1 * v;
            }
            s_fscale[f] = (float)sc;
        }
    }
    __syncthreads();   // the only CTA-wide barrier

    // spans blockIdx.x, blockIdx.x + gridDim.x, ... (all full: the host sends the ragged tail to the generic kernel)
    const long long gstride = gridDim.x;
    const int n_local = (int)((p.n_spans - blockIdx.x + gstride - 1) / gstride);

    if (warp == kConsumerWarps) {
        // ---- loader warp: loads n_stages spans ahead, retires spans in order (bulk store of the gradients) ----
        if (lane != 0) return;
        auto issue_load = [&](int it) {
            const long long span = blockIdx.x + (long long)it * gstride;
            const int s = it % n_stages;
            float* dst_t = ring + (size_t)s * stage_floats;
            constexpr uint32_t bt = (uint32_t)(t_floats * 4), bq = (uint32_t)(q_floats * 4);
            mbar_arrive_expect_tx(&full_bar[s], bt + bq);
            bulk_g2s(dst_t, p.yt + span * (long long)(TP * ST_T), bt, &full_bar[s]);
            bulk_g2s(dst_t + t_floats, p.yp + span * (long long)(TP * ST_P), bq, &full_bar[s]);
        };
        for (int it = 0; it < n_stages && it < n_local; ++it) issue_load(it);
        for (int j = 0; j < n_local; ++j) {
            const int s = j % n_stages;
            mbar_wait(&done_bar[s], (uint32_t)(j / n_stages) & 1u);
            const long long span = blockIdx.x + (long long)j * gstride;
            bulk_s2g(bp.grad + span * (long long)(TP * CP), ring + (size_t)s * stage_floats + t_floats + q_floats, (uint32_t)(g_floats * 4));
            // the stage retired one span ago is reloaded once its gradient store has finished reading shared memory
            if (j >= 1 && j - 1 + n_stages < n_local) {
                asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                issue_load(j - 1 + n_stages);
            }
        }
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // stores complete before the CTA retires
        return;
    }

    // ---- consumer warps ----
    const float fscale = s_focal_scale;
    const int wch = p.wch;
    int s = 0;
    uint32_t parity = 0;
    for (int it = 0; it < n_local; ++it) {
        mbar_wait(&full_bar[s], parity);
        const float* __restrict__ T = ring + (size_t)s * stage_floats + tid * ST_T;
        const float* __restrict__ Q = ring + (size_t)s * stage_floats + t_floats + tid * ST_P;
        float* __restrict__ G = ring + (size_t)s * stage_floats + t_floats + q_floats + tid * CP;
        const float w = wch >= 0 ? T[wch] : 1.0f;
        float ymax = 0.f;
#pragma unroll
        for (int c = 0; c < HM; ++c) ymax = fmaxf(ymax, T[c]);
        if (ymax >= 1.0f) {
            bwd_peak_pixel(p, bp, T, Q, G, w, fscale, s_fscale);
        } else {
            const float ws = w * fscale;
#pragma unroll
            for (int c = 0; c < HM; ++c) {
                const float y = T[c], q = Q[c];
                const float u = 1.0f - q;
                const float l = log_one_minus(q);                                        // log(clip(1 - q, .01, .99))
                const float dl = (u >= 0.01f && u <= 0.99f) ? __fdividef(-1.0f, u) : 0.0f;   // clip_by_value passes grad inside
                const float t = 1.0f - y, t2 = t * t;
                G[c] = (-(t2 * t2) * (2.0f * q * l + (q * q) * dl)) * ws;                // d neg_loss / d q, loss.py:43-48
            }
#pragma unroll
            for (int c = HM; c < CP; ++c) G[c] = 0.f;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the bulk engine
        __syncwarp();
        if (lane == 0) mbar_arrive(&done_bar[s]);
        if (++s == n_stages) {
            s = 0;
            parity ^= 1u;
        }
    }
}

template <int HM, int ST_T, int ST_P, int CP>
int launch_bwd_fast(BwdParams& bp, long long n_full_spans, cudaStream_t st) {
    LossParams& p = bp.fwd;
    p.TP = kBwdThreads;
    p.n_spans = n_full_spans;
    const size_t stage_bytes = (size_t)kBwdThreads * (ST_T + ST_P + CP) * 4;
    int n_stages = (int)((size_t)(110 * 1024) / stage_bytes);   // two CTAs per SM
    if (n_stages > kBwdStages) n_stages = kBwdStages;
    if (n_stages < 2) return CVM_ERR_ARG;
    p.n_stages = n_stages;
    const size_t smem = (size_t)n_stages * stage_bytes;
    const int grid = loss_grid(p.n_spans, smem);
    CVM_SMEM_ATTR_ONCE((loss_bwd_fast_kernel<HM, ST_T, ST_P, CP>), smem);
    loss_bwd_fast_kernel<HM, ST_T, ST_P, CP><<<grid, kBwdThreads + 32, smem, st>>>(bp);
    CVM_CHECK_LAUNCH("loss_bwd_fast_kernel");
    return CVM_OK;
}

}  // namespace

namespace {
int loss_bwd_impl(const cvm_layout* L, const float* y_true, int y_true_stride, const float* y_pred, int y_pred_stride,
                  long long n_pixels, const double* partials, const float* upstream, float* grad_pred, bool force_generic,
                  void* stream) {
    int rc = check_layout_for_loss(L, y_true_stride, y_pred_stride);
    if (rc != CVM_OK) return rc;
    CVM_CHECK_ARG(y_true && y_pred && partials && grad_pred, "NULL pointer argument");
    CVM_CHECK_ARG(n_pixels >= 0, "n_pixels < 0");
    if (n_pixels == 0) return CVM_OK;
    BwdParams bp;
    memset(&bp, 0, sizeof(bp));
    LossParams& p = bp.fwd;
    // ring (3 stages) + two gradient buffers within ~100 KB
    int TP = 256;
    const int per_px = kStages * (y_true_stride + y_pred_stride) + 2 * L->Cp;
    while (TP > 16 && (size_t)TP * per_px * 4 > 100 * 1024) TP >>= 1;
    const size_t smem = (size_t)TP * per_px * 4;
    p.TP = TP;
    p.yt = y_true;
    p.yp = y_pred;
    p.n_pixels = n_pixels;
    p.n_spans = (n_pixels + TP - 1) / TP;
    p.st_t = y_true_stride;
    p.st_p = y_pred_stride;
    p.hm = L->hm;
    p.wch = L->Ct - 1;
    p.inv_hm = 1.0f / (float)L->hm;
    p.fa = L->focal_a;
    p.fb = L->focal_b;
    p.a_is2 = (L->focal_a == 2.0f);
    p.b_is4 = (L->focal_b == 4.0f);
    p.n_fields = L->n_fields;
    for (int f = 0; f < L->n_fields; ++f) {
        p.f_off[f] = L->field_off[f];
        p.f_size[f] = L->field_size[f];
        p.f_kind[f] = L->field_kind[f];
        bp.f_post[f] = L->field_post[f];
        bp.f_weight[f] = L->field_weight[f];
    }
    p.use_bulk = cvm_aligned16(y_true) && cvm_aligned16(y_pred);
    bp.partials = partials;
    bp.upstream = upstream;
    bp.grad = grad_pred;
    bp.Cp = L->Cp;
    bp.grad_bulk = cvm_aligned16(grad_pred);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    // fast path: compile-time layout, everything 16-byte aligned, alpha = 2 / beta = 4; it takes the full spans, the generic
    // kernel the ragged tail (or everything)
    long long done_pixels = 0;
    if (p.use_bulk && bp.grad_bulk && p.a_is2 && p.b_is4 && !force_generic) {
        const long long n_full = n_pixels / kBwdThreads;
        BwdParams fb = bp;
        int frc = CVM_ERR_ARG;
        if (n_full > 0) {
            if (L->hm == 10 && y_true_stride == 15 && y_pred_stride == 14 && L->Cp == 14) frc = launch_bwd_fast<10, 15, 14, 14>(fb, n_full, st);
            else if (L->hm == 10 && y_true_stride == 17 && y_pred_stride == 16 && L->Cp == 16) frc = launch_bwd_fast<10, 17, 16, 16>(fb, n_full, st);
            else if (L->hm == 1 && y_true_stride == 16 && y_pred_stride == 15 && L->Cp == 15) frc = launch_bwd_fast<1, 16, 15, 15>(fb, n_full, st);
        }
        if (frc == CVM_OK) done_pixels = n_full * kBwdThreads;
        else if (frc == CVM_ERR_CUDA) return frc;
    }
    if (done_pixels < n_pixels) {   // generic kernel on [done_pixels, n_pixels)
        p.yt = y_true + done_pixels * y_true_stride;
        p.yp = y_pred + done_pixels * y_pred_stride;
        bp.grad = grad_pred + done_pixels * L->Cp;
        p.n_pixels = n_pixels - done_pixels;
        p.n_spans = (p.n_pixels + TP - 1) / TP;
        p.use_bulk = cvm_aligned16(p.yt) && cvm_aligned16(p.yp);
        bp.grad_bulk = cvm_aligned16(bp.grad);
        const int grid = loss_grid(p.n_spans, smem);
        CVM_SMEM_ATTR_ONCE((loss_bwd_kernel), smem);
        loss_bwd_kernel<<<grid, kThreads, smem, st>>>(bp);
        CVM_CHECK_LAUNCH("loss_bwd_kernel");
    }
    return CVM_OK;
}
}  // namespace

extern "C" int cvm_loss_bwd(const cvm_layout* L, const float* y_true, int y_true_stride, const float* y_pred,
                            int y_pred_stride, long long n_pixels, const double* partials, const float* upstream,
                            float* grad_pred, void* stream) {
    return loss_bwd_impl(L, y_true, y_true_stride, y_pred, y_pred_stride, n_pixels, partials, upstream, grad_pred, false, stream);
}

// the same gradient from the layout-generic kernel only (what every layout without a compile-time instantiation gets); the
// tests compare the two on the BASELINE shapes
extern "C" int cvm_loss_bwd_generic(const cvm_layout* L, const float* y_true, int y_true_stride, const float* y_pred,
                                    int y_pred_stride, long long n_pixels, const double* partials, const float* upstream,
                                    float* grad_pred, void* stream) {
    return loss_bwd_impl(L, y_true, y_true_stride, y_pred, y_pred_stride, n_pixels, partials, upstream, grad_pred, true, stream);
}
