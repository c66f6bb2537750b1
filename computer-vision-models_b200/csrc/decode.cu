// Output decode, canonical CenterNet form (north_star): 3x3 peak NMS fused with an exact
// per-image top-K, then gather of the regression heads and box assembly (box math of the reference's
// models/centernet/post_processing.py:43-52 and common/utils/image.py:22-28, in fp32 like NumPy does it).
//
// Bound: HBM reads.  Algorithmic bytes per image: 4*H*W*pred_stride, read exactly once.
//
// Kernel 1 (decode_scan_kernel).  The batch is one flat list of steps (one or two granules of 480 consecutive pixels of one
// image, all channels: a contiguous piece of y_pred).  The list is cut into gridDim.x equal contiguous ranges, one persistent CTA
// per SM and range, so every SM streams the same number of bytes.  A loader warp feeds a shared-memory ring of granules
// with 1-D bulk async copies (TMA engine: UBLKCP + mbarrier, full/empty barrier pair per slot),
// several granules ahead.  Fifteen scanner warps walk the steps WITHOUT any CTA-wide barrier: per step every lane owns
// one pixel per granule of the step, takes the maximum over its heatmap channels (vector shared loads, no bank conflicts)
// and compares it ONCE with the CTA's running threshold score.  A step is scanned when the W + 1 pixels after it have
// arrived too, so a pixel that reaches the threshold is tested right away: the 3x3 test reads the eight neighbours out
// of the ring (W + 1 pixels of history and lookahead stay resident; for maps too wide for that they are read from global
// memory).  Peaks are appended to a shared-memory candidate buffer as 64-bit keys (score bits << 32 | ~flat index): a total order with no ties, equal
// to (score desc, flat index asc).  A score histogram of the appended peaks raises the threshold (one warp scans it
// between steps).  Only two rare events gather the scanner warps on a named barrier: the buffer passing its mark (exact
// radix select, keeps the best K) and the end of an image, where the CTA writes its <= K best keys of that image
// ("segment") to the workspace and starts over.
// Kernel 2 (decode_merge_kernel): one CTA per image merges the segments of the CTAs that touched it (same radix select),
// rank-sorts the K winners, fills a short tail with score-0 entries in flat-index order (tf.nn.top_k semantics on the
// masked map), gathers r_offset / fullbox / track_offset at the peaks and assembles boxes.
#include <limits.h>
#include <stdlib.h>

#include <type_traits>

#include "common.cuh"

namespace {

constexpr int kScanThreads = 480;  // scanner threads: one pixel per thread and step (15 warps + the loader warp = 4 warps per
                                   // SM sub-partition, which leaves 128 registers per thread; a 17th warp caps them at 96 and spills)
constexpr int kScanWarps = kScanThreads / 32;
constexpr int kThreads = kScanThreads + 32;   // + the loader warp
constexpr int kMergeThreads = 256;
constexpr int kMaxSlots = 32;      // ring depth (granules) the barrier arrays can hold
constexpr int kSlack = 1536;       // buffer entries beyond K before the (rare) fallback compaction
constexpr int kSegKeys = 512;      // keys a segment may publish without an exact select (the merge kernel selects anyway)
#ifndef CVM_SCORE_SHIFT
#define CVM_SCORE_SHIFT 19
#endif
constexpr int kScoreShift = CVM_SCORE_SHIFT;    // score histogram: bin = float bits >> 19 (sign 0, 8 exponent bits, 4 mantissa bits)
constexpr int kScoreBins = 1 << (31 - kScoreShift);
constexpr int kMaxK = 1024;
constexpr int kMergeCap = 4096;    // merge kernel: keys buffered before an intermediate select
constexpr int kSmemBudget = 227 * 1024 - 1024;

struct DecodeParams {
    const float* yp;
    int stride, H, W, hm, K;
    int HW;                       // pixels per image
    int T;                        // pixels per ring granule (<= kScanThreads)
    int gps;                      // granules per step (pixels per scanner lane and step)
    int spi, gpi;                 // steps / granules per image
    long long n_steps;            // B * spi
    int S;                        // ring slots (granules)
    int ring_nb;                  // 1: neighbours are read from the ring; 0: from global memory
    int halo_g;                   // granules of history / lookahead that hold W + 1 pixels (0 without ring_nb)
    int gran_floats;              // T * stride
    int cap;                      // candidate buffer entries
    int compact_at;               // compact when a candidate lands at or beyond this position
    int max_segs;                 // images one CTA range can touch
    int seg_keys;                 // keys a segment can publish (>= K)
    unsigned long long* keys;     // [grid][max_segs][seg_keys]
    int* counts;                  // [grid][max_segs]
    unsigned int thr0_bits;       // initial threshold score bits (1 = smallest positive float; experiments only raise it)
    float inv_W;                  // 1 / W
    int rescan_step;              // appends after which the score histogram is scanned again (test_hits)
    int dbg;                      // experiment knob CVM_DECODE_DBG: 1 = no pixel scan, 2 = no hit test, 4 = also rescan the histogram at every step
};

// Fixed-size head of the dynamic shared memory block; the ring, the candidate buffer, the select scratch and the score
// histogram follow (see smem_layout()).
struct SharedHead {
    uint64_t full_bar[kMaxSlots];   // loader -> scanners: granule arrived (tx bytes, or a plain arrive without the bulk engine)
    uint64_t empty_bar[kMaxSlots];  // scanners -> loader: all scanner warps are done with the granule
    unsigned int hist[256];       // radix select
    int misc[4];
    // control word, read by every scanner warp once per step with ONE 64-bit shared load (a warp-wide broadcast, so
    // all lanes of a warp always see the same pair)
    unsigned int thr_bits;        // running threshold score (float bits), raised by the histogram scan
    int compact_flag;             // an append landed at/after the compaction mark
    int count;                    // candidates buffered
    int scanned;                  // `count` when the histogram was last scanned
    int maxbin;                   // highest score bin seen so far
    int seg_done;                 // scanner warps that finished the current segment
    unsigned long long thr;       // K-th key of the last exact select
#ifdef CVM_WATCHDOG
    int dbg_state[16];
#endif
};
constexpr int kHeadBytes = (sizeof(SharedHead) + 127) & ~127;

// The dynamic shared memory block, declared at file scope so that every device function addresses it in the shared
// state space (LDS/STS/ATOMS); passing views of it through structs or pointers degrades them to generic accesses.
extern __shared__ __align__(128) unsigned char g_smem[];

__device__ __forceinline__ SharedHead* sm_head() { return reinterpret_cast<SharedHead*>(g_smem); }
#ifdef CVM_WATCHDOG
__device__ __forceinline__ int* sm_head_raw();
#endif
#ifdef CVM_WATCHDOG
__device__ __forceinline__ int* sm_head_raw() { return sm_head()->dbg_state; }
#endif
// [S][gran_floats] (+ 4 floats of slack for vector over-reads of the last pixel)
__device__ __forceinline__ float* sm_ring() { return reinterpret_cast<float*>(g_smem + kHeadBytes); }
// [cap] candidate keys
__device__ __forceinline__ unsigned long long* sm_cand(const DecodeParams& p) {
    return reinterpret_cast<unsigned long long*>(g_smem + kHeadBytes + ((size_t)p.S * p.gran_floats + 4) * 4);
}
// [K] select scratch
__device__ __forceinline__ unsigned long long* sm_keep(const DecodeParams& p) { return sm_cand(p) + p.cap; }
// [kScoreBins] score histogram of the buffered candidates
__device__ __forceinline__ unsigned int* sm_shist(const DecodeParams& p) {
    return reinterpret_cast<unsigned int*>(sm_keep(p) + p.K);
}

size_t smem_bytes(int S, int gran_floats, int cap, int K) {
    return (size_t)kHeadBytes + ((size_t)S * gran_floats + 4) * 4 + (size_t)cap * 8 + (size_t)K * 8 + (size_t)kScoreBins * 4;
}

#ifdef CVM_WATCHDOG
#define DBG_STATE(code) do { if ((threadIdx.x & 31) == 0) ((volatile int*)sm_head_raw())[threadIdx.x >> 5] = (code); } while (0)
#else
#define DBG_STATE(code) do { } while (0)
#endif

// -DCVM_DECODE_STATS: per-warp cycle / event counters for tuning (tools/decode_stats.py); never in the shipped build
#ifdef CVM_DECODE_STATS
__device__ unsigned long long g_decode_stats[16];
#define STAT_DECL long long st_t0 = 0; unsigned long long st_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0}
#define STAT_BEGIN() (st_t0 = clock64())
#define STAT_END(i) (st_acc[i] += (unsigned long long)(clock64() - st_t0))
#define STAT_ADD(i, v) (st_acc[i] += (unsigned long long)(v))
#define STAT_FLUSH() do { if ((threadIdx.x & 31) == 0) { for (int k_ = 0; k_ < 8; ++k_) atomicAdd(&g_decode_stats[k_], st_acc[k_]); } } while (0)
#else
#define STAT_DECL
#define STAT_BEGIN() ((void)0)
#define STAT_END(i) ((void)0)
#define STAT_ADD(i, v) ((void)0)
#define STAT_FLUSH() ((void)0)
#endif
#ifdef CVM_DECODE_STATS
#define TH_MARK(i) do { const long long t_ = clock64(); if ((threadIdx.x & 31) == 0) { atomicAdd(&g_decode_stats[i], (unsigned long long)(t_ - th_t)); if ((i) == 12) atomicAdd(&g_decode_stats[8], 1ull); } th_t = clock64(); } while (0)
#define TH_START() long long th_t = clock64()
#else
#define TH_MARK(i) ((void)0)
#define TH_START() ((void)0)
#endif

// named barrier over the first `nt` threads of the CTA (the scanner warps; the loader warp never joins)
__device__ __forceinline__ void group_sync(int nt) { asm volatile("bar.sync 1, %0;" ::"r"(nt) : "memory"); }


__device__ __noinline__ void gather(const DecodeParams& p, bool at_segment_end, int seg);

// ---- exact top-K select on distinct 64-bit keys held in shared memory -------------------------------------------------
// On return keys[0..K) hold the K largest (unordered), *thr is the K-th largest key.  n > K required.  Called by the
// first nt threads of the CTA.
__device__ __noinline__ void select_topk(unsigned long long* keys, int n, int K, unsigned int* hist, unsigned long long* keep,
                                         int* s_misc /* [4] */, unsigned long long* thr_out, int nt) {
    const int tid = threadIdx.x;
    unsigned long long prefix = 0ull, mask = 0ull;
    int need = K;
    for (int shift = 56; shift >= 0; shift -= 8) {
        if (tid < 256) hist[tid] = 0;
        group_sync(nt);
        for (int i = tid; i < n; i += nt) {
            const unsigned long long k = keys[i];
            if ((k & mask) == prefix) atomicAdd(&hist[(unsigned)(k >> shift) & 255u], 1u);
        }
        group_sync(nt);
        if (tid < 32) {
            // lane l owns bins [255-8l-7, 255-8l], scanned from the top
            unsigned int loc[8], sum = 0;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                loc[j] = hist[255 - (tid * 8 + j)];
                sum += loc[j];
            }
            unsigned int incl = sum;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned int t = __shfl_up_sync(0xffffffffu, incl, o);
                if (tid >= o) incl += t;
            }
            unsigned int above = incl - sum;  // keys in strictly higher bins than this lane's range
            if (above < (unsigned)need && incl >= (unsigned)need) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    if (above < (unsigned)need && above + loc[j] >= (unsigned)need) {
                        s_misc[0] = 255 - (tid * 8 + j);   // selected digit
                        s_misc[1] = need - (int)above;     // still needed inside that bin
                        s_misc[2] = (int)loc[j];           // population of that bin
                    }
                    above += loc[j];
                }
            }
        }
        group_sync(nt);
        const int digit = s_misc[0];
        need = s_misc[1];
        const int pop = s_misc[2];
        prefix |= (unsigned long long)digit << shift;
        mask |= 0xFFull << shift;
        group_sync(nt);  // s_misc is rewritten next pass
        if (pop == need) break;  // the whole bin is selected: every key >= prefix (low bits zero) is kept
    }
    const unsigned long long T = prefix;
    if (tid == 0) s_misc[3] = 0;
    group_sync(nt);
    for (int i = tid; i < n; i += nt) {
        const unsigned long long k = keys[i];
        if (k >= T) keep[atomicAdd(&s_misc[3], 1)] = k;
    }
    group_sync(nt);
    for (int i = tid; i < K; i += nt) keys[i] = keep[i];
    // T is the K-th largest key with (after an early exit) its undecided low bits zeroed: a valid, at most marginally
    // weaker, lower bound for every later candidate
    if (tid == 0) *thr_out = T;
    group_sync(nt);
}

// Threshold scan (one warp): the largest score bin tb with at least K buffered candidates in bins >= tb.  Every counted
// entry is a real peak of this image, so a later candidate below the lower edge of tb cannot make the top K: the edge
// becomes the new threshold.  Counts read while other warps append are at worst too low, which only makes the bound
// conservative.  No buffer traffic, no CTA-wide sync.
__device__ __forceinline__ void scan_threshold(SharedHead* h, const unsigned int* shist, int K, int lane) {
    int n = 0, top = 0;
    if (lane == 0) {   // one lane decides (the words change under our feet), the warp follows
        n = *(volatile const int*)&h->count;
        top = *(volatile const int*)&h->maxbin;
        if (n == *(volatile const int*)&h->scanned || n < K) n = -1;
    }
    n = __shfl_sync(0xffffffffu, n, 0);
    top = __shfl_sync(0xffffffffu, top, 0);
    if (n < 0) return;
    unsigned above = 0;
    for (int base = top; base >= 0; base -= 32) {
        const int bin = base - lane;
        const unsigned v = bin >= 0 ? ((volatile const unsigned int*)shist)[bin] : 0u;
        unsigned incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        const unsigned hit = __ballot_sync(0xffffffffu, above + incl >= (unsigned)K);
        if (hit) {
            const int tb = base - (__ffs(hit) - 1);
            if (lane == 0) {
                atomicMax(&h->thr_bits, (unsigned)tb << kScoreShift);   // several warps may scan at once
                h->scanned = n;
            }
            return;
        }
        above += __shfl_sync(0xffffffffu, incl, 31);
    }
    if (lane == 0) h->scanned = n;
}

// Sum of `v` over the first nt threads of the CTA (one shared atomic per warp); `slot` is scratch.
__device__ __forceinline__ int __syncthreads_count_scan(int v, int nt, int* slot) {
    if (threadIdx.x == 0) *slot = 0;
    group_sync(nt);
    v = __reduce_add_sync(0xffffffffu, v);
    if ((threadIdx.x & 31) == 0 && v) atomicAdd(slot, v);
    group_sync(nt);
    const int total = *(volatile int*)slot;
    group_sync(nt);
    return total;
}

// Fallback when the candidate buffer passed the mark: exact select, then the histogram is rebuilt from the K survivors.
__device__ __noinline__ void compact_buffer(const DecodeParams& p) {
    const int tid = threadIdx.x, nt = kScanThreads, K = p.K;
    SharedHead* h = sm_head();
    unsigned long long* cand = sm_cand(p);
    unsigned int* shist = sm_shist(p);
    select_topk(cand, h->count, K, h->hist, sm_keep(p), h->misc, &h->thr, nt);
    for (int i = tid; i < kScoreBins; i += nt) shist[i] = 0u;
    group_sync(nt);
    for (int i = tid; i < K; i += nt) atomicAdd(&shist[(unsigned)(cand[i] >> (32 + kScoreShift))], 1u);
    if (tid == 0) {
        h->count = K;
        h->scanned = K;
        const unsigned bits = (unsigned)(h->thr >> 32) & ~((1u << kScoreShift) - 1u);
        if (bits > h->thr_bits) h->thr_bits = bits;
        h->compact_flag = 0;
    }
    group_sync(nt);
}

// End of a (CTA, image) segment: publish the candidates that still reach the running threshold (the merge kernel does
// the exact select; only a segment with more than kSegKeys of them selects its best K first), reset the selection
// state.  All scanner threads, count frozen.
__device__ __noinline__ void flush_segment(const DecodeParams& p, int seg) {
    const int tid = threadIdx.x, nt = kScanThreads;
    SharedHead* h = sm_head();
    unsigned long long* cand = sm_cand(p);
    unsigned int* shist = sm_shist(p);
    const size_t o = (size_t)blockIdx.x * p.max_segs + seg;
    unsigned long long* out = p.keys + o * p.seg_keys;
    int n = h->count;
    const unsigned thr_bits = h->thr_bits;
    int above = 0;
    for (int i = tid; i < n; i += nt) above += (unsigned)(cand[i] >> 32) >= thr_bits;
    above = __syncthreads_count_scan(above, nt, &h->misc[0]);
    if (above > p.seg_keys) {   // rare: heavy ties at the threshold score
        select_topk(cand, n, p.K, h->hist, sm_keep(p), h->misc, &h->thr, nt);
        n = p.K;
        for (int i = tid; i < n; i += nt) out[i] = cand[i];
    } else {
        if (tid == 0) h->misc[1] = 0;
        group_sync(nt);
        for (int i = tid; i < n; i += nt) {
            const unsigned long long k = cand[i];
            if ((unsigned)(k >> 32) >= thr_bits) out[atomicAdd(&h->misc[1], 1)] = k;
        }
        n = above;
    }
    for (int i = tid; i < kScoreBins; i += nt) shist[i] = 0u;
    group_sync(nt);
    if (tid == 0) {
        p.counts[o] = n;
        h->count = 0;
        h->thr = 0ull;
        h->thr_bits = p.thr0_bits;
        h->scanned = 0;
        h->maxbin = 0;
        h->compact_flag = 0;
        h->seg_done = 0;
    }
    group_sync(nt);
}

__device__ __forceinline__ uint2 load_ctrl(const SharedHead* h) {
    uint2 v;
    asm volatile("ld.volatile.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(smem_u32(&h->thr_bits)) : "memory");
    return v;
}

// One lane's pixel that reached the threshold and waits for its 3x3 test.
// channel masks: 32 bits when the compile-time channel count allows it (registers are tight in the scan kernel)
template <int HM>
using mask_t = typename std::conditional<(HM > 0 && HM <= 32), unsigned int, unsigned long long>::type;
template <typename M>
__device__ __forceinline__ int lowest_bit(M m) { return sizeof(M) == 4 ? __ffs((int)m) - 1 : __ffsll((long long)m) - 1; }

template <int HM>
struct Hit {
    mask_t<HM> mask;           // heatmap channels whose score reached the threshold (0: no hit)
    int q;                     // pixel index inside the image
    int rp;                    // ring position (in pixels) of the pixel
    float vmax;                // largest heatmap score of the pixel
};

// Offsets (in floats) of channel 0 of the eight neighbours of a hit pixel: into the ring (W + 1 pixels of halo are
// resident) or relative to the pixel in global memory; INT_MIN outside the map.  Interior pixels whose neighbourhood does
// not wrap around the ring (almost all) take the short way.
template <int HM>
__device__ __forceinline__ void neighbour_offsets(const DecodeParams& p, const Hit<HM>& hit, int (&nb)[8]) {
    const int W = p.W, stride = p.stride, ring_px = p.S * p.T;
    // y = q / W without an integer division (exact after one correction step for any q < 2^31)
    const int q = hit.q;
    int y = __float2int_rz(__fmul_rn((float)q, p.inv_W)), x = q - y * W;
    if (x < 0) {
        --y;
        x += W;
    } else if (x >= W) {
        ++y;
        x -= W;
    }
    const bool interior = y > 0 && y < p.H - 1 && x > 0 && x < W - 1;
    const bool flat = !p.ring_nb || (hit.rp - W - 1 >= 0 && hit.rp + W + 1 < ring_px);
    if (interior && flat) {
        const int base = p.ring_nb ? hit.rp * stride : 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int j = k < 4 ? k : k + 1;
            nb[k] = base + ((j / 3 - 1) * W + (j % 3 - 1)) * stride;
        }
    } else {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int j = k < 4 ? k : k + 1;
            const int dy = j / 3 - 1, dx = j % 3 - 1;
            const int yy = y + dy, xx = x + dx;
            int off = INT_MIN;
            if (yy >= 0 && yy < p.H && xx >= 0 && xx < W) {
                const int dq = dy * W + dx;
                if (p.ring_nb) {
                    int r = hit.rp + dq;
                    if (r < 0) r += ring_px;
                    if (r >= ring_px) r -= ring_px;
                    off = r * stride;
                } else {
                    off = dq * stride;
                }
            }
            nb[k] = off;
        }
    }
}

// is channel ch of the hit pixel a peak (its value equals its 3x3 maximum; plateaus are all kept) that still reaches the
// threshold (it may have risen since the scan; it is > 0, so score-0 entries never get here)?
template <int HM>
__device__ __forceinline__ bool is_peak(const DecodeParams& p, const float* ring, const float* g_px, const Hit<HM>& hit,
                                        const int (&nb)[8], int ch, float thr_f, float& v) {
    v = ring[hit.rp * p.stride + ch];
    if (!(v >= thr_f)) return false;
    float m = v;
    if (p.ring_nb) {
#pragma unroll
        for (int k = 0; k < 8; ++k)
            if (nb[k] != INT_MIN) m = fmaxf(m, ring[nb[k] + ch]);
    } else {
#pragma unroll
        for (int k = 0; k < 8; ++k)
            if (nb[k] != INT_MIN) m = fmaxf(m, g_px[nb[k] + ch]);
    }
    return m == v;
}

// bits of the heatmap channels of one pixel whose score reaches the threshold (called for the few pixels whose maximum does)
template <int HM>
__device__ __forceinline__ mask_t<HM> channel_mask(const float* px, int hm, float thr_f) {
    mask_t<HM> m = 0;
    if (HM > 0) {
        unsigned lo = 0u;   // HM <= 32 in every instantiation
#pragma unroll
        for (int c = 0; c < HM; ++c) lo |= (px[c] >= thr_f ? 1u : 0u) << c;
        m = lo;
    } else {
        for (int c = 0; c < hm; ++c)
            if (px[c] >= thr_f) m |= (mask_t<HM>)1 << c;
    }
    return m;
}

// Append the peaks found in one round (peak[k] of this lane's k-th pixel, channel ch[k], score v[k]) to the candidate buffer:
// one shared atomic per warp, keys placed by ballot rank.  Warp-uniform call (pm[k] = ballot of peak[k], total = their sum > 0).
// Ends at a safe point: joins a pending compaction, or rescans the score histogram when enough keys have come in.
template <int HM, int NH>
__device__ __forceinline__ void append_peaks(const DecodeParams& p, const Hit<HM> (&hit)[NH], const bool (&peak)[NH], const float (&v)[NH],
                                             const int (&ch)[NH], const unsigned (&pm)[NH], unsigned total) {
    SharedHead* h = sm_head();
    unsigned long long* cand = sm_cand(p);
    unsigned int* shist = sm_shist(p);
    const int lane = threadIdx.x & 31;
    unsigned base = 0;
    int rescan = 0;
    TH_START();
    if (lane == 0) {
        base = (unsigned)atomicAdd(&h->count, (int)total);
        const int cnt = (int)(base + total);
        rescan = cnt >= p.K && cnt - *(volatile int*)&h->scanned >= p.rescan_step;
    }
    base = __shfl_sync(0xffffffffu, base, 0);
    rescan = __shfl_sync(0xffffffffu, rescan, 0);
    TH_MARK(12);
#pragma unroll
    for (int k = 0; k < NH; ++k) {
        if (peak[k]) {
            const unsigned pos = base + __popc(pm[k] & ((1u << lane) - 1u));
            const unsigned bits = __float_as_uint(v[k]);
            const unsigned flat = (unsigned)hit[k].q * (unsigned)p.hm + (unsigned)ch[k];
            cand[pos] = ((unsigned long long)bits << 32) | (unsigned long long)(0xFFFFFFFFu - flat);
            if (pos >= (unsigned)p.compact_at) *(volatile int*)&h->compact_flag = 1;
            const unsigned bin = bits >> kScoreShift;
            atomicAdd(&shist[bin], 1u);
            atomicMax(&h->maxbin, (int)bin);   // (no result used: nothing waits for it)
        }
        base += __popc(pm[k]);
    }
    // Safe point after every batch of appends: once the buffer has passed its mark, a warp adds at most these <= 32 * NH
    // keys before it joins the compaction, which bounds the buffer (see plan_decode).
    TH_MARK(13);
    __syncwarp();
    const bool must_gather = __any_sync(0xffffffffu, load_ctrl(h).y != 0);
    TH_MARK(14);
    if (must_gather) gather(p, false, 0);
    else if (rescan) scan_threshold(h, shist, p.K, lane);
    TH_MARK(15);
}

// Exact 3x3 test of the pending hits of a warp (NH pixels per lane) and append of the peaks.  Called by ALL lanes of the
// warp (lanes without a hit pass mask 0).  Every round each pixel tests ONE of its pending channels (the channels differ
// between lanes, the control flow does not: the loop condition is a vote, the ballots and the aggregated append - one
// shared atomic per warp and round - are warp-uniform); the NH pixels of a lane are tested side by side (independent load
// chains).  While the segment has no threshold yet (`thr_f` is the initial one) the first round takes each pixel's
// LARGEST channel: the first K peaks found that way are high ones, and the threshold they give prunes most of the rest.
// The warp that pushes the count K / 2 past the last histogram scan rescans, and the pending channels are re-filtered
// whenever the threshold has moved.  img: image of the pixels; thr_f: in/out, the warp's current threshold.
template <int HM, int NH>
__device__ __forceinline__ int test_hits(const DecodeParams& p, long long img, const Hit<HM> (&hit)[NH], float& thr_f) {
    using M = mask_t<HM>;
    M rem[NH], any = 0;
#pragma unroll
    for (int k = 0; k < NH; ++k) {
        rem[k] = hit[k].mask;
        any |= rem[k];
    }
    if (!__any_sync(0xffffffffu, any != 0)) return 0;
    TH_START();
    SharedHead* h = sm_head();
    const float* ring = sm_ring();
    // ---- fast path (nearly every call once the segment has a threshold): each hit pixel of the warp has ONE pending
    //      channel and its 3x3 neighbourhood is inside the map.  Straight-line code: the nine
    //      loads are issued while the row / column of the pixel is still being worked out, a vote confirms that nobody was
    //      at the border, and three quarters of the calls end at the ballot because no hit was a peak. ----
    if (p.ring_nb && __float_as_uint(thr_f) != p.thr0_bits) {
        const int W = p.W, stride = p.stride, ring_px = p.S * p.T;
        bool cheap = true;
#pragma unroll
        for (int k = 0; k < NH; ++k)
            if (rem[k]) cheap = cheap && (rem[k] & (rem[k] - 1)) == 0;
        if (__all_sync(0xffffffffu, cheap)) {
            bool peak[NH], border = false;
            float v[NH];
            int ch[NH];
            unsigned pm[NH], total = 0;
#pragma unroll
            for (int k = 0; k < NH; ++k) {
                peak[k] = false;
                v[k] = 0.f;
                ch[k] = 0;
                if (rem[k]) {
                    ch[k] = lowest_bit(rem[k]);
                    // ring positions of the three rows' centre pixels and of their left / right neighbours (the ring is a
                    // circular buffer of pixels: every step may wrap)
                    const int mid = hit[k].rp;
                    int up = mid - W, dn = mid + W;
                    up += up < 0 ? ring_px : 0;
                    dn -= dn >= ring_px ? ring_px : 0;
                    const int ul = up == 0 ? ring_px - 1 : up - 1, ur = up == ring_px - 1 ? 0 : up + 1;
                    const int ml = mid == 0 ? ring_px - 1 : mid - 1, mr = mid == ring_px - 1 ? 0 : mid + 1;
                    const int dl = dn == 0 ? ring_px - 1 : dn - 1, dr = dn == ring_px - 1 ? 0 : dn + 1;
                    const float* c = ring + ch[k];
                    const float a0 = c[ul * stride], a1 = c[up * stride], a2 = c[ur * stride], a3 = c[ml * stride], a4 = c[mr * stride],
                                a5 = c[dl * stride], a6 = c[dn * stride], a7 = c[dr * stride];
                    v[k] = c[mid * stride];
                    const int q = hit[k].q;
                    int y = __float2int_rz(__fmul_rn((float)q, p.inv_W)), x = q - y * W;   // exact after one correction step
                    if (x < 0) {
                        --y;
                        x += W;
                    } else if (x >= W) {
                        ++y;
                        x -= W;
                    }
                    border = border || !(y > 0 && y < p.H - 1 && x > 0 && x < W - 1);
                    const float m = fmaxf(fmaxf(fmaxf(a0, a1), fmaxf(a2, a3)), fmaxf(fmaxf(a4, a5), fmaxf(a6, a7)));
                    peak[k] = v[k] >= thr_f && m <= v[k];
                }
            }
            if (!__any_sync(0xffffffffu, border)) {
#pragma unroll
                for (int k = 0; k < NH; ++k) {
                    pm[k] = __ballot_sync(0xffffffffu, peak[k]);
                    total += __popc(pm[k]);
                }
                if (total) append_peaks<HM, NH>(p, hit, peak, v, ch, pm, total);
                return 1;
            }
        }
    }
    const int lane = threadIdx.x & 31;
    (void)lane;
    (void)h;
    int nb[NH][8];
    const float* g_px[NH];
    M pref[NH];
    int rounds = 0;
    const bool young = __float_as_uint(thr_f) == p.thr0_bits;
#pragma unroll
    for (int k = 0; k < NH; ++k) {
        g_px[k] = p.yp + ((size_t)img * p.HW + (size_t)hit[k].q) * p.stride;
        pref[k] = 0;
        if (hit[k].mask) {
            neighbour_offsets(p, hit[k], nb[k]);
            if (young) {
                const M top = channel_mask<HM>(ring + (size_t)hit[k].rp * p.stride, p.hm, hit[k].vmax) & hit[k].mask;
                pref[k] = top & ((M)0 - top);
            }
        }
    }
    do {
        bool peak[NH];
        float v[NH];
        int ch[NH];
        unsigned pm[NH], total = 0;
#pragma unroll
        for (int k = 0; k < NH; ++k) {
            v[k] = 0.f;
            const M sel = pref[k] ? pref[k] : (rem[k] & ((M)0 - rem[k]));
            pref[k] = 0;
            ch[k] = sel ? lowest_bit(sel) : 0;
            peak[k] = sel != 0 && is_peak(p, ring, g_px[k], hit[k], nb[k], ch[k], thr_f, v[k]);
            rem[k] &= ~sel;
        }
#pragma unroll
        for (int k = 0; k < NH; ++k) {
            pm[k] = __ballot_sync(0xffffffffu, peak[k]);
            total += __popc(pm[k]);
        }
        if (total) append_peaks<HM, NH>(p, hit, peak, v, ch, pm, total);
        // the threshold may have moved (our rescan, another warp's, a compaction): drop the pending channels below it
        const unsigned now_bits = __shfl_sync(0xffffffffu, load_ctrl(h).x, 0);
        if (now_bits > __float_as_uint(thr_f)) {
            thr_f = __uint_as_float(now_bits);
#pragma unroll
            for (int k = 0; k < NH; ++k)
                if (rem[k]) rem[k] &= channel_mask<HM>(ring + (size_t)hit[k].rp * p.stride, p.hm, thr_f);
        }
        any = 0;
#pragma unroll
        for (int k = 0; k < NH; ++k) any |= rem[k];
        ++rounds;
    } while (__any_sync(0xffffffffu, any != 0));
    return rounds;   // statistics only
}

// maximum over the HM leading floats of one pixel; STRIDE > 0: compile-time layout, widest aligned vector loads
template <int STRIDE, int HM>
__device__ __forceinline__ float pixel_max(const float* px, int hm) {
    float m = __int_as_float(0xff800000);
    if (STRIDE == 0) {
        for (int c = 0; c < hm; ++c) m = fmaxf(m, px[c]);
    } else if (STRIDE % 4 == 0) {
#pragma unroll
        for (int c = 0; c < HM; c += 4) {
            const float4 t = *reinterpret_cast<const float4*>(px + c);
            m = fmaxf(m, t.x);
            if (c + 1 < HM) m = fmaxf(m, t.y);
            if (c + 2 < HM) m = fmaxf(m, t.z);
            if (c + 3 < HM) m = fmaxf(m, t.w);
        }
    } else if (STRIDE % 2 == 0) {
#pragma unroll
        for (int c = 0; c < HM; c += 2) {
            const float2 t = *reinterpret_cast<const float2*>(px + c);
            m = fmaxf(m, t.x);
            if (c + 1 < HM) m = fmaxf(m, t.y);
        }
    } else {
#pragma unroll
        for (int c = 0; c < HM; ++c) m = fmaxf(m, px[c]);
    }
    return m;
}

// Gathering of the scanner warps (called warp-uniformly at a step boundary).  A warp comes here when it saw the
// compaction flag, or (at_segment_end) when it has finished the segment and wants the flush; whoever arrives waits for
// all of them, then everybody takes the same decisions from state that cannot change while all are paused.
__device__ __noinline__ void gather(const DecodeParams& p, bool at_segment_end, int seg) {
    SharedHead* h = sm_head();
#ifdef CVM_DECODE_STATS
    const long long g_t0 = clock64();
    struct GatherTimer {
        long long t0;
        bool seg;
        __device__ ~GatherTimer() {
            if ((threadIdx.x & 31) == 0) {
                if (seg) atomicAdd(&g_decode_stats[10], (unsigned long long)(clock64() - t0));
                if (seg) atomicAdd(&g_decode_stats[11], 1ull);
            }
        }
    } g_timer{g_t0, at_segment_end};
#endif
    for (;;) {
        DBG_STATE(at_segment_end ? 21 : 20);
        group_sync(kScanThreads);   // everybody paused: no appends in flight
        DBG_STATE(22);
        const int flag = *(volatile int*)&h->compact_flag;
        const int done = *(volatile int*)&h->seg_done;
        group_sync(kScanThreads);
        if (done == kScanWarps) {   // all warps finished the segment: publish it (selects when more than K are buffered)
            flush_segment(p, seg);
            return;
        }
        DBG_STATE(23);
        if (flag) compact_buffer(p);        // clears the flag
        DBG_STATE(24);
        if (!at_segment_end) return;        // back to scanning; a warp waiting for the flush keeps gathering
    }
}



// Per-thread (warp-uniform) pipeline state of a scanner warp; lives in registers.
struct ScanState {
    float thr_f;          // threshold score this warp currently compares with
    int waited, w_slot;   // granules [0, waited) have arrived; slot of granule `waited`
    uint32_t w_parity;    // parity of the full-barrier phase of granule `waited`
    int released, r_slot; // granules [0, released) were handed back by this warp
};

// Wait until granules [0, upto) have arrived.
__device__ __forceinline__ void wait_until(const DecodeParams& p, ScanState& z, int upto) {
    SharedHead* const h = sm_head();
    while (z.waited < upto) {
        // A warp blocked on data is at a safe point (it is not appending): it must join a compaction gather, or the
        // warps behind it (whose releases the loader needs) would wait for it forever.  The votes keep the warp
        // converged: the barrier can complete, and the flag can change, between two lanes' polls.
        while (!__all_sync(0xffffffffu, mbar_try_wait(&h->full_bar[z.w_slot], z.w_parity))) {
            if (__any_sync(0xffffffffu, load_ctrl(h).y != 0)) {
                gather(p, false, 0);
                z.thr_f = __uint_as_float(load_ctrl(h).x);
            }
        }
        ++z.waited;
        if (++z.w_slot == p.S) {
            z.w_slot = 0;
            z.w_parity ^= 1u;
        }
    }
}

// Hand granules [released, upto) back to the loader (one arrive per warp and granule).
__device__ __forceinline__ void release_until(const DecodeParams& p, ScanState& z, int upto) {
    if (z.released < upto) {
        SharedHead* const h = sm_head();
        __syncwarp();   // every lane is done reading them
        while (z.released < upto) {
            if ((threadIdx.x & 31) == 0) mbar_arrive(&h->empty_bar[z.r_slot]);
            ++z.released;
            if (++z.r_slot == p.S) z.r_slot = 0;
        }
    }
}

// STRIDE/HM = 0: runtime pixel stride / heatmap channel count.  GPS: granules (= pixels per lane) per step.  BULK: granules
// arrive by bulk async copy (needs 16-byte aligned granules); otherwise the loader warp copies them with plain loads.
template <int STRIDE, int HM, int GPS, bool BULK>
__global__ void __launch_bounds__(kThreads, 1) decode_scan_kernel(const __grid_constant__ DecodeParams p) {
    SharedHead* const h = sm_head();
    float* const ring = sm_ring();
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int S = p.S, T = p.T, spi = p.spi, gpi = p.gpi, HW = p.HW, hg = p.halo_g;
    const int stride = STRIDE ? STRIDE : p.stride;

    // this CTA's contiguous range of the flat step list (a step = GPS consecutive granules of one image); granule `seq`
    // of the CTA is flat granule g_first + seq (every image has gpi granules, so flat granule indices are contiguous)
    const long long G = gridDim.x, g = blockIdx.x;
    const long long s0 = g * p.n_steps / G, s1 = (g + 1) * p.n_steps / G;
    if (s0 >= s1) return;
    const long long img0 = s0 / spi, imgL = (s1 - 1) / spi;
    const int st0 = (int)(s0 - img0 * spi), stL = (int)((s1 - 1) - imgL * spi);
    const int n_local = (int)(s1 - s0);
    const int lead = min(hg, st0 * GPS);                       // history granules before the first step (same image only)
    const int endL = min(gpi, (stL + 1) * GPS);                // first granule after the last step
    const int tail = min(hg, gpi - endL);                      // lookahead granules after the last step (same image only)
    const int n_load = (int)((imgL - img0) * gpi + endL + tail - (st0 * GPS - lead));

    if (tid == 0) {
        for (int k = 0; k < S; ++k) {
            mbar_init(&h->full_bar[k], 1);
            mbar_init(&h->empty_bar[k], kScanWarps);
        }
        mbar_fence_init();
        h->count = 0;
        h->thr = 0ull;
        h->thr_bits = p.thr0_bits;   // smallest positive float: "score > 0" and "score >= threshold" in one compare
        h->scanned = 0;
        h->maxbin = 0;
        h->compact_flag = 0;
        h->seg_done = 0;
    }
    for (int k = tid; k < kScoreBins; k += kThreads) sm_shist(p)[k] = 0u;
    __syncthreads();   // the only CTA-wide barrier: from here on the loader warp and the scanner warps run on their own

    if (warp == kScanWarps) {
        // ---- loader warp: granule `seq` goes to slot seq % S once all scanner warps have released its previous tenant ----
        int slot = 0, gi = st0 * GPS - lead;
        long long l_img = img0;
        uint32_t e_parity = 0;   // parity of the empty-barrier phase that frees a slot for its next tenant
        for (int seq = 0; seq < n_load; ++seq) {
#ifdef CVM_WATCHDOG
            if (seq >= S) {
                long long spins = 0;
                while (!mbar_try_wait(&h->empty_bar[slot], e_parity)) {
                    if (++spins > 16000000LL) {
                        if (lane == 0) {
                            printf("loader stuck: cta %d seq %d n_load %d slot %d flag %d seg_done %d count %d\n", (int)blockIdx.x, seq,
                                   n_load, slot, h->compact_flag, h->seg_done, h->count);
                            volatile int* d = h->dbg_state;
                            printf("  states cta %d: %d %d %d %d %d %d %d %d %d %d %d %d %d %d %d %d\n", (int)blockIdx.x, d[0], d[1], d[2], d[3],
                                   d[4], d[5], d[6], d[7], d[8], d[9], d[10], d[11], d[12], d[13], d[14], d[15]);
                        }
                        __trap();
                    }
                }
            }
#else
            if (seq >= S) mbar_wait(&h->empty_bar[slot], e_parity);
#endif
            const int npx = min(T, HW - gi * T);
            const float* src = p.yp + ((size_t)l_img * HW + (size_t)gi * T) * stride;
            float* dst = ring + (size_t)slot * p.gran_floats;
            if (BULK) {
                if (lane == 0) {
                    const uint32_t bytes = (uint32_t)npx * (uint32_t)stride * 4u;
                    mbar_arrive_expect_tx(&h->full_bar[slot], bytes);
                    bulk_g2s(dst, src, bytes, &h->full_bar[slot]);
                }
            } else {
                const int nf = npx * stride;
                for (int k = lane; k < nf; k += 32) dst[k] = src[k];
                __syncwarp();
                if (lane == 0) mbar_arrive(&h->full_bar[slot]);   // release: the stores above are visible to the waiters
            }
            if (++slot == S) {
                slot = 0;
                if (seq >= S) e_parity ^= 1u;
            }
            if (++gi == gpi) {
                gi = 0;
                ++l_img;
            }
        }
        return;
    }

    // ---- scanner warps ----
    ScanState z;
    z.thr_f = __uint_as_float(p.thr0_bits);
    z.waited = z.w_slot = z.released = z.r_slot = 0;
    z.w_parity = 0u;

    STAT_DECL;
    long long img = img0;
    int st = st0;
    int gseq = lead;         // sequence number of the first granule of the current step
    int cur_slot = lead;     // its ring slot (lead <= halo_g < S)
    for (int i = 0; i < n_local; ++i) {
        const int g0 = st * GPS;                                 // first granule of the step inside the image
        const int gc = min(GPS, gpi - g0);                        // granules in this step (the last step of an image may be short)
        // last granule of this image that this CTA fetches
        const int seq_last = min(gseq + (gpi - 1 - g0), n_load - 1);
        // this step's granules and their lookahead (W + 1 pixels past the end of the step): a pixel is scanned and, if it
        // reaches the threshold, tested in the same step
        DBG_STATE(100 + i * 1000);
        STAT_BEGIN();
        wait_until(p, z, min(gseq + gc - 1 + hg, seq_last) + 1);
        STAT_END(0);
        STAT_BEGIN();
        DBG_STATE(101 + i * 1000);
        Hit<HM> cur[GPS];
#pragma unroll
        for (int k = 0; k < GPS; ++k) {
            int sl = cur_slot + k;
            if (sl >= S) sl -= S;
            const int q0 = (g0 + k) * T;
            cur[k].mask = 0;
            cur[k].q = q0 + tid;
            cur[k].rp = sl * T + tid;   // ring position of this lane's pixel
            if (k < gc && tid < min(T, HW - q0) && !(p.dbg & 1)) {
                const float* px = ring + (size_t)cur[k].rp * stride;
                cur[k].vmax = pixel_max<STRIDE, HM>(px, p.hm);
                if (cur[k].vmax >= z.thr_f) cur[k].mask = channel_mask<HM>(px, p.hm, z.thr_f);
            }
        }
        DBG_STATE(102 + i * 1000);
        STAT_END(1);
#ifdef CVM_DECODE_STATS
        for (int k = 0; k < GPS; ++k) {
            STAT_ADD(4, __popc(__ballot_sync(0xffffffffu, cur[k].mask != 0)));                 // pixel hits
        }
        STAT_ADD(6, 1);                                                                          // warp-steps
#endif
        STAT_BEGIN();
#ifdef CVM_DECODE_STATS
        const bool was_young = __float_as_uint(z.thr_f) == p.thr0_bits;
        const int n_rounds = test_hits<HM, GPS>(p, img, cur, z.thr_f);
        STAT_ADD(5, n_rounds);
        if (was_young) { STAT_END(7); } else { STAT_END(2); }
#else
        if (!(p.dbg & 2)) test_hits<HM, GPS>(p, img, cur, z.thr_f);
#endif
        STAT_BEGIN();
        const bool segment_end = (st + 1 == spi) || (i + 1 == n_local);
        const bool image_end = st + 1 == spi;
        // the next step needs halo_g granules of history before its first granule: the rest is dead
        release_until(p, z, segment_end ? (image_end ? gseq + gc : n_load) : gseq + gc - hg);
        __syncwarp();   // converged: the control word below is one broadcast load, the same pair for all lanes
        const uint2 ctrl = load_ctrl(h);
        z.thr_f = __uint_as_float(ctrl.x);   // stale values are still valid bounds
        DBG_STATE(103 + i * 1000);
        if (!segment_end) {
            // (the histogram is rescanned by the appending warps, test_hits; a rescan by one warp at every step on top of that
            // cost 4 % - an out-of-line call spills live registers, and local memory is an L2 round trip here)
            if (warp == 1 && (p.dbg & 4)) scan_threshold(h, sm_shist(p), p.K, lane);
            if (__any_sync(0xffffffffu, ctrl.y != 0)) {   // a vote: the decision to gather must be warp-uniform
                gather(p, false, 0);
                z.thr_f = __uint_as_float(load_ctrl(h).x);
            }
            ++st;
        } else {
            if (lane == 0) atomicAdd(&h->seg_done, 1);
            gather(p, true, (int)(img - img0));
            z.thr_f = __uint_as_float(p.thr0_bits);
            if (image_end) {
                st = 0;
                ++img;
            }
        }
        gseq += gc;
        cur_slot += gc;
        if (cur_slot >= S) cur_slot -= S;
        STAT_END(3);
    }
    STAT_FLUSH();
}

struct MergeParams {
    const float* yp;
    int stride, H, W, hm, K;
    int spi, grid, max_segs, seg_keys;
    long long n_steps;
    int off_roff, off_box, off_track;
    int off_class, nb_classes;    // Profile R (hm == 1 + class-logit field): cls = first argmax of the logits
    float R;
    const cvm_roi* rois;
    const unsigned long long* keys;
    const int* counts;
    float* scores;
    int32_t* cls;
    long long* flat;
    float* centers;
    float* boxes;
    float* track;
};

__global__ void __launch_bounds__(kMergeThreads) decode_merge_kernel(const MergeParams p) {
    __shared__ unsigned int hist[256];
    __shared__ int s_misc[4];
    __shared__ unsigned long long s_thr;

    const int tid = threadIdx.x, b = blockIdx.x, K = p.K;
    unsigned long long* const all = reinterpret_cast<unsigned long long*>(g_smem);   // [kMergeCap + K]
    unsigned long long* const keep = all + kMergeCap + K;                              // [K]
    unsigned long long* const sorted = keep + K;                                       // [K]

    // the scan CTAs whose step range [g*n/G, (g+1)*n/G) overlaps this image's steps [lo, hi)
    const long long n_ch = p.n_steps, G = p.grid;
    const long long lo = (long long)b * p.spi, hi = lo + p.spi;
    long long g = lo * G / n_ch;
    while (g + 1 < G && (g + 1) * n_ch / G <= lo) ++g;
    int n = 0;
    for (; g < G; ++g) {
        const long long s0 = g * n_ch / G, s1 = (g + 1) * n_ch / G;
        if (s0 >= hi) break;
        if (s1 <= s0) continue;
        const int seg = (int)(b - s0 / p.spi);
        const size_t o = (size_t)g * p.max_segs + seg;
        const int cnt = p.counts[o];
        if (n + cnt > kMergeCap) {   // cnt <= seg_keys <= kMergeCap / 4, so n > K here
            group_sync(kMergeThreads);
            select_topk(all, n, K, hist, keep, s_misc, &s_thr, kMergeThreads);
            n = K;
        }
        const unsigned long long* src = p.keys + o * p.seg_keys;
        for (int i = tid; i < cnt; i += kMergeThreads) all[n + i] = src[i];
        n += cnt;
    }
    group_sync(kMergeThreads);
    if (n > K) {
        select_topk(all, n, K, hist, keep, s_misc, &s_thr, kMergeThreads);
        n = K;
    }
    // rank sort (keys are distinct): descending
    for (int i = tid; i < n; i += kMergeThreads) {
        const unsigned long long k = all[i];
        int rank = 0;
        for (int j = 0; j < n; ++j) rank += all[j] > k;
        sorted[rank] = k;
    }
    group_sync(kMergeThreads);

    const int hm = p.hm, W = p.W;
    const long long n_total = (long long)p.H * W * hm;
    const int K_eff = (long long)K < n_total ? K : (int)n_total;
    const cvm_roi roi = p.rois ? p.rois[b] : cvm_roi{1.0f, 0.0f, 0.0f, 0.0f};
    for (int i = tid; i < K; i += kMergeThreads) {
        const size_t o = (size_t)b * K + i;
        float score = 0.f;
        long long fl = -1;
        if (i < n) {
            const unsigned long long k = sorted[i];
            score = __uint_as_float((unsigned)(k >> 32));
            fl = (long long)(0xFFFFFFFFu - (unsigned)(k & 0xFFFFFFFFull));
        } else if (i < K_eff) {
            // tail: the (i-n)-th flat index that is not among the n positive peaks, in ascending order
            const long long j = i - n;
            long long f = j;
            for (;;) {
                int c = 0;
                for (int q = 0; q < n; ++q) c += (long long)(0xFFFFFFFFu - (unsigned)(sorted[q] & 0xFFFFFFFFull)) <= f;
                if (j + c == f) break;
                f = j + c;
            }
            fl = f;
        }
        p.scores[o] = score;
        p.flat[o] = fl;
        float cx = 0.f, cy = 0.f, bw = 0.f, bh = 0.f, tx = 0.f, ty = 0.f;
        int cl = -1;
        if (fl >= 0) {
            const long long pix = fl / hm;
            cl = (int)(fl - pix * hm);
            const int y = (int)(pix / W), x = (int)(pix - (long long)y * W);
            const float* px = p.yp + (((size_t)b * p.H + y) * W + x) * p.stride;
            if (hm == 1 && p.off_class >= 0) {   // np.argmax of the class logits: first maximum (post_processing.py:39-41)
                float best = px[p.off_class];
                cl = 0;
                for (int k = 1; k < p.nb_classes; ++k) {
                    const float v = px[p.off_class + k];
                    if (v > best) {
                        best = v;
                        cl = k;
                    }
                }
            }
            float dx = 0.f, dy = 0.f, w = 0.f, h = 0.f;
            if (p.off_roff >= 0) {
                dx = px[p.off_roff];
                dy = px[p.off_roff + 1];
            }
            if (p.off_box >= 0) {
                w = px[p.off_box];
                h = px[p.off_box + 1];
            }
            // post_processing.py:44-52 + image.py:22-28, every op rounded to fp32 (no FMA contraction)
            cx = __fsub_rn(__fmul_rn(roi.inv_scale, __fmul_rn(__fadd_rn((float)x, dx), p.R)), roi.off_left);
            cy = __fsub_rn(__fmul_rn(roi.inv_scale, __fmul_rn(__fadd_rn((float)y, dy), p.R)), roi.off_top);
            bw = __fmul_rn(w, roi.inv_scale);
            bh = __fmul_rn(h, roi.inv_scale);
            if (p.off_track >= 0) {
                // the tracking OFFSET in roi coordinates: centers + track = predicted centre in the previous frame
                // (targets: prev centre - centre, centertracker/processor.py:82-89), what cvm_track_associate consumes
                tx = __fmul_rn(px[p.off_track], roi.inv_scale);
                ty = __fmul_rn(px[p.off_track + 1], roi.inv_scale);
            }
        }
        p.cls[o] = cl;
        p.centers[o * 2 + 0] = cx;
        p.centers[o * 2 + 1] = cy;
        p.boxes[o * 4 + 0] = __fsub_rn(cx, __fmul_rn(bw, 0.5f));
        p.boxes[o * 4 + 1] = __fsub_rn(cy, __fmul_rn(bh, 0.5f));
        p.boxes[o * 4 + 2] = bw;
        p.boxes[o * 4 + 3] = bh;
        if (p.track) {
            p.track[o * 2 + 0] = tx;
            p.track[o * 2 + 1] = ty;
        }
    }
}


struct Plan {
    int T, gps, spi, gpi, S, ring_nb, halo_g, gran_floats, cap, compact_at, grid, max_segs, seg_keys;
    long long n_steps;
    size_t smem_scan, smem_merge, ws_keys, ws_total;
};

int env_int(const char* name, int dflt) {
    const char* e = getenv(name);
    return e ? atoi(e) : dflt;
}

int plan_decode(const cvm_layout* L, int stride, int B, int K, Plan* t) {
    const int W = L->W;
    const long long HW = (long long)L->H * W;
    // granule: one pixel per scanner thread, fewer when the image is smaller; a step is 2 granules when the ring has room
    // (the per-step bookkeeping and the latency of the hit tests are then paid once per 2 pixels of every lane)
    int T = env_int("CVM_DECODE_T", kScanThreads);
    if (T > kScanThreads || T < 32 || T % 32) T = kScanThreads;
    if (HW < T) T = (int)((HW + 31) / 32 * 32);
    t->compact_at = K + env_int("CVM_DECODE_SLACK", kSlack);
    t->cap = t->compact_at + 1 + 2 * kScanThreads;   // every warp stops within 32 appends per pixel of a lane of the mark (test_hits)
    const size_t fixed = smem_bytes(0, 0, t->cap, K);
    int gps_want = env_int("CVM_DECODE_GPS", 2);
    if (gps_want < 1 || gps_want > 2 || HW <= T) gps_want = 1;
    for (;; T -= 32) {
        if (T < 32) return CVM_ERR_ARG;
        const size_t gran_bytes = (size_t)T * stride * 4;
        if (fixed + 4 * gran_bytes > (size_t)kSmemBudget) continue;
        int S = (int)(((size_t)kSmemBudget - fixed) / gran_bytes);
        if (S > kMaxSlots) S = kMaxSlots;
        // ring mode keeps halo_g + gps + halo_g granules resident (history, the step, its lookahead); global-neighbour mode
        // keeps the step; both want at least two, better three, more in flight
        const int hg = (W + 1 + T - 1) / T;
        t->gps = 0;
        for (int gps = gps_want; gps >= 1 && !t->gps; --gps) {
            const int need = hg + gps + hg + 2;
            if (S >= need) {
                t->gps = gps;
                t->ring_nb = 1;
                t->halo_g = hg;
                t->S = S > need ? need + 1 : need;
            }
        }
        if (!t->gps) {
            t->gps = S >= 4 ? gps_want : 1;
            t->ring_nb = 0;
            t->halo_g = 0;
            t->S = S < t->gps + 2 ? S : t->gps + 2;
        }
        t->gran_floats = T * stride;
        break;
    }
    {
        const int v = env_int("CVM_DECODE_S", 0);   // experiment knob: ring depth
        const int need = t->ring_nb ? 2 * t->halo_g + t->gps + 1 : t->gps + 1;
        if (v >= need && v <= kMaxSlots && smem_bytes(v, t->gran_floats, t->cap, K) <= (size_t)kSmemBudget) t->S = v;
    }
    t->T = T;
    t->gpi = (int)((HW + T - 1) / T);
    t->spi = (t->gpi + t->gps - 1) / t->gps;
    t->n_steps = (long long)B * t->spi;
    // CVM_DECODE_SPARE_SMS (experiment knob): SMs left free, so that the small all-reduce of the loss partials, issued just
    // before the decode by a data-parallel caller, finds an SM and runs beside it (2 GPUs: 0.634 -> 0.632 ms per step)
    long long grid = cvm_num_sms() - env_int("CVM_DECODE_SPARE_SMS", 0);
    if (grid > t->n_steps) grid = t->n_steps;
    if (grid < 1) grid = 1;
    t->grid = (int)grid;
    const long long len = (t->n_steps + grid - 1) / grid;
    t->max_segs = (int)((len - 1) / t->spi + 2);
    t->smem_scan = smem_bytes(t->S, t->gran_floats, t->cap, K);
    t->smem_merge = ((size_t)kMergeCap + 3 * (size_t)K) * 8;
    t->seg_keys = K > kSegKeys ? K : kSegKeys;
    t->ws_keys = (size_t)grid * t->max_segs * t->seg_keys * 8;
    t->ws_total = t->ws_keys + (size_t)grid * t->max_segs * 4;
    return CVM_OK;
}

int check_decode_args(const cvm_layout* L, int stride, int B, int K) {
    CVM_CHECK_ARG(L != nullptr, "layout is NULL");
    CVM_CHECK_ARG(L->H > 0 && L->W > 0 && B >= 0, "bad shape");
    CVM_CHECK_ARG(L->hm >= 1 && L->hm <= 64 && stride >= L->Cp && L->Cp >= L->hm, "bad channel layout");
    CVM_CHECK_ARG(stride <= 256, "pixel stride above 256 floats is not supported");
    CVM_CHECK_ARG(K >= 1 && K <= kMaxK, "K=%d outside [1,%d]", K, kMaxK);
    CVM_CHECK_ARG((long long)L->H * L->W * L->hm < 0xFFFFFFFFLL, "H*W*hm must fit in 32 bits");
    return CVM_OK;
}

template <int STRIDE, int HM, int GPS, bool BULK>
int launch_scan_one(const DecodeParams& p, const Plan& t, cudaStream_t st) {
    CVM_CHECK_CUDA(cudaFuncSetAttribute(decode_scan_kernel<STRIDE, HM, GPS, BULK>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)t.smem_scan));
    decode_scan_kernel<STRIDE, HM, GPS, BULK><<<t.grid, kThreads, t.smem_scan, st>>>(p);
    CVM_CHECK_LAUNCH("decode_scan_kernel");
    return CVM_OK;
}

template <int STRIDE, int HM>
int launch_scan(const DecodeParams& p, const Plan& t, bool bulk, cudaStream_t st) {
    if (t.gps == 2) return bulk ? launch_scan_one<STRIDE, HM, 2, true>(p, t, st) : launch_scan_one<STRIDE, HM, 2, false>(p, t, st);
    return bulk ? launch_scan_one<STRIDE, HM, 1, true>(p, t, st) : launch_scan_one<STRIDE, HM, 1, false>(p, t, st);
}

}  // namespace

extern "C" size_t cvm_decode_topk_workspace_bytes(const cvm_layout* L, int pred_stride, int B, int K) {
    if (check_decode_args(L, pred_stride, B, K) != CVM_OK) return 0;
    Plan t;
    if (plan_decode(L, pred_stride, B > 0 ? B : 1, K, &t) != CVM_OK) return 0;
    return t.ws_total;
}

extern "C" int cvm_decode_topk(const cvm_layout* L, const float* y_pred, int pred_stride, int B, int K,
                               const cvm_roi* rois, float* scores, int32_t* cls, long long* flat, float* centers,
                               float* boxes, float* track, void* ws, size_t ws_bytes, void* stream) {
    int rc = check_decode_args(L, pred_stride, B, K);
    if (rc != CVM_OK) return rc;
    CVM_CHECK_ARG(y_pred && scores && cls && flat && centers && boxes && ws, "NULL pointer argument");
    if (B == 0) return CVM_OK;
    Plan t;
    rc = plan_decode(L, pred_stride, B, K, &t);
    CVM_CHECK_ARG(rc == CVM_OK, "no tiling for H=%d W=%d hm=%d stride=%d K=%d", L->H, L->W, L->hm, pred_stride, K);
    if (ws_bytes < t.ws_total) {
        cvm_set_error("workspace too small: %zu < %zu", ws_bytes, t.ws_total);
        return CVM_ERR_WS;
    }
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    DecodeParams p;
    memset(&p, 0, sizeof(p));
    p.yp = y_pred;
    p.stride = pred_stride;
    p.H = L->H;
    p.W = L->W;
    p.hm = L->hm;
    p.K = K;
    p.HW = L->H * L->W;
    p.T = t.T;
    p.gps = t.gps;
    p.spi = t.spi;
    p.gpi = t.gpi;
    p.n_steps = t.n_steps;
    p.S = t.S;
    p.ring_nb = t.ring_nb;
    p.halo_g = t.halo_g;
    p.gran_floats = t.gran_floats;
    p.cap = t.cap;
    p.compact_at = t.compact_at;
    p.max_segs = t.max_segs;
    p.seg_keys = t.seg_keys;
    p.keys = static_cast<unsigned long long*>(ws);
    p.counts = reinterpret_cast<int*>(static_cast<unsigned char*>(ws) + t.ws_keys);
    p.thr0_bits = 1u;
    p.inv_W = 1.0f / (float)L->W;
    p.dbg = env_int("CVM_DECODE_DBG", 0);
    p.rescan_step = env_int("CVM_DECODE_RESCAN", K / 2 > 8 ? K / 2 : 8);
    if (const char* e = getenv("CVM_DECODE_THR0")) {   // experiment knob (results are wrong when set): start threshold
        const float f = (float)atof(e);
        memcpy(&p.thr0_bits, &f, 4);
    }

    // the bulk-copy engine needs 16-byte granules: base pointer aligned and every image a whole number of them (full
    // granules are Pg*stride*4 bytes with Pg % 32 == 0, the partial last granule of an image then ends on one too)
    const bool bulk = cvm_aligned16(y_pred) && (((long long)p.HW * pred_stride) % 4 == 0);
    if (pred_stride == 14 && L->hm == 10) rc = launch_scan<14, 10>(p, t, bulk, st);        // CenterNet, 10 classes
    else if (pred_stride == 16 && L->hm == 10) rc = launch_scan<16, 10>(p, t, bulk, st);   // CenterTracker
    else if (pred_stride == 20 && L->hm == 10) rc = launch_scan<20, 10>(p, t, bulk, st);   // multitask head
    else rc = launch_scan<0, 0>(p, t, bulk, st);
    if (rc != CVM_OK) return rc;

    MergeParams m;
    memset(&m, 0, sizeof(m));
    m.yp = y_pred;
    m.stride = pred_stride;
    m.H = L->H;
    m.W = L->W;
    m.hm = L->hm;
    m.K = K;
    m.spi = t.spi;
    m.grid = t.grid;
    m.max_segs = t.max_segs;
    m.seg_keys = t.seg_keys;
    m.n_steps = t.n_steps;
    m.off_roff = L->off_roff;
    m.off_box = L->off_box;
    m.off_track = track ? L->off_track : -1;
    m.off_class = L->off_class;
    m.nb_classes = L->nb_classes;
    m.R = (float)L->R;
    m.rois = rois;
    m.keys = p.keys;
    m.counts = p.counts;
    m.scores = scores;
    m.cls = cls;
    m.flat = flat;
    m.centers = centers;
    m.boxes = boxes;
    m.track = track;
    CVM_CHECK_CUDA(cudaFuncSetAttribute(decode_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)t.smem_merge));
    decode_merge_kernel<<<B, kMergeThreads, t.smem_merge, st>>>(m);
    CVM_CHECK_LAUNCH("decode_merge_kernel");
    return CVM_OK;
}

#ifdef CVM_DECODE_STATS
extern "C" int cvm_decode_stats(unsigned long long* out16, int reset) {
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(out16, g_decode_stats, sizeof(unsigned long long) * 16);
    if (reset) {
        unsigned long long z[16] = {0};
        cudaMemcpyToSymbol(g_decode_stats, z, sizeof(z));
    }
    return 0;
}
#endif
