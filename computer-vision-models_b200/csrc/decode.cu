// Output decode, canonical CenterNet form (north_star): 3x3 peak NMS fused with an exact
// per-image top-K, then gather of the regression heads and box assembly (box math of the reference's
// models/centernet/post_processing.py:43-52 and common/utils/image.py:22-28, in fp32 like NumPy does it).  Optionally
// the per-pixel argmax of a semseg slice of the SAME tensor (to_3channel's class pick, common/utils/image.py:72-100;
// multitask head, models/multitask/loss.py:21-32) is taken in the same pass, so a multitask output is read once.
//
// Bound: HBM reads.  Algorithmic bytes per image: 4*H*W*pred_stride, read exactly once (+ H*W bytes written for the
// class map when the argmax is fused).
//
// Kernel 1 (decode_scan_kernel), one persistent CTA per SM, three kinds of warps that never meet on a CTA-wide barrier:
//   * loader warp: the batch is one flat list of granules (T consecutive pixels of one image, all channels: a contiguous
//     piece of y_pred) cut into gridDim.x equal contiguous ranges; the loader streams its range through a shared-memory
//     ring with 1-D bulk async copies (TMA engine: UBLKCP + mbarrier, full/empty barrier pair per slot).
//   * scanner warps: warp w owns granules w, w + kScanWarps, ... of the range.  Per pixel: the maximum over the heatmap
//     channels (vector shared loads) and ONE compare with the segment's running threshold score; the few pixels that
//     reach it are pushed as (pixel, channel mask) records into small shared-memory queues.  Nothing else: the slot goes
//     back to the loader as soon as the granule has been scanned, so every ring slot is available for data in flight.
//   * tester warps: drain the queues.  A record is tested exactly (value equal to the maximum of its 3x3 neighbourhood in
//     its channel, plateaus all kept); the neighbours are read from global memory - they were streamed through L2 a few
//     microseconds earlier, so these are L2 hits, a few hundred bytes per test, and any map width works the same way.
//     Peaks are appended to a shared candidate buffer as 64-bit keys (score bits << 32 | ~flat index): a total order
//     without ties, equal to (score desc, flat index asc).  A score histogram of the appended peaks raises the threshold
//     (the lower edge of the highest bin with K peaks at or above it), which the scanners pick up at their next step.
//   A (CTA, image) pair is a "segment" with its own top-K: when every scanner warp has left the image, the testers publish
//   the segment's surviving keys to the workspace and reset the selection state.  Bootstrap of a segment: the warp that
//   owns its first granule pushes only each pixel's best channel for the better half of the pixels ("hints"); the K-th best
//   of the peaks found among them is the first threshold, then that granule is scanned again for everything else above it.
// Kernel 2 (decode_merge_kernel): one CTA per image merges the segments of the CTAs that touched it (exact radix select),
// rank-sorts the K winners, fills a short tail with score-0 entries in flat-index order (tf.nn.top_k semantics on the
// masked map), gathers r_offset / fullbox / track_offset (/ class logits) at the peaks and assembles boxes.
#include <limits.h>
#include <stdlib.h>

#include <type_traits>

#include "common.cuh"

namespace {

#ifndef CVM_TEST_WARPS
#define CVM_TEST_WARPS 6
#endif
#ifndef CVM_POLL_NS
#define CVM_POLL_NS 32
#endif
constexpr int kTestWarps = CVM_TEST_WARPS;      // threads [0, kTestThreads): tester warps (they share named barrier 1)
constexpr int kScanWarps = 8;
constexpr int kTestThreads = kTestWarps * 32;
constexpr int kLoaderWarp = kTestWarps + kScanWarps;
constexpr int kThreads = (kLoaderWarp + 1) * 32;
constexpr int kMergeThreads = 256;
constexpr int kMergeTab = 192;     // scan CTAs whose image ranges travel in the merge kernel's parameters
constexpr int kMaxSlots = kScanWarps;   // ring depth (granules): one slot per active scanner warp
constexpr int kCtr = 2 * kMaxSlots;
constexpr int kMaxT = 1024;        // pixels per granule (a scanner lane keeps one hit bit per pixel of its granule: 32 x 32)
constexpr int kGranBytes = 32768;  // target granule size
#ifndef CVM_UNROLL
#define CVM_UNROLL 4
#endif
constexpr int kUnroll = CVM_UNROLL;   // pixels per scanner lane and iteration
constexpr int kQShift = 7;
constexpr int kQCap = 1 << kQShift;   // records per queue (one queue per tester warp and segment parity)
#ifndef CVM_SLACK
#define CVM_SLACK 1024
#endif
constexpr int kSlack = CVM_SLACK;       // buffer entries beyond K before the (rare) fallback compaction
constexpr int kSegKeys = 512;      // keys a segment may publish without an exact select (the merge kernel selects anyway)
constexpr int kScoreShift = 19;    // score histogram: bin = float bits >> 19 (sign 0, 8 exponent bits, 4 mantissa bits)
constexpr int kScoreBins = 1 << (31 - kScoreShift);
constexpr int kMaxK = 1024;
constexpr int kMergeCap = 4096;    // merge kernel: keys buffered before an intermediate select
constexpr int kSmemBudget = 227 * 1024 - 1024;
constexpr unsigned kFull = 0xffffffffu;

// queue records: lo = lap bit | marker bit | pixel index (or marker code), hi = kind of record
constexpr unsigned kRecMarker = 1u << 30, kRecPixel = (1u << 30) - 1u;
constexpr unsigned kKindAll = 1u;      // test every channel of the pixel that reaches the threshold
constexpr unsigned kKindHint = 2u;     // hint: test the pixel's best channel only
constexpr unsigned kKindRest = 3u;     // second pass over a hinted granule: like kKindAll without the channel a hint covered
constexpr unsigned kMarkEnd = 1u, kMarkHintEnd = 2u;

// what a scan CTA leaves per (CTA, image) segment next to its keys
struct SegMeta {
    int count;                    // keys published
    unsigned int thr_bits;        // score bits every published key reaches
    unsigned int verified;        // 1: at least K peaks of the segment reach thr_bits, or the segment was bootstrapped exactly;
                                  // 0: thr_bits is still the provisional threshold the segment started with
    unsigned int pad;
};

struct DecodeParams {
    const float* yp;
    int stride, H, W, hm, K;
    int HW;                       // pixels per image
    int T;                        // pixels per ring granule
    int gpi;                      // granules per image
    long long n_gran;             // B * gpi
    int S;                        // ring slots (granules) = active scanner warps: slot w belongs to scanner warp w, which
                                  // therefore waits for its full-barrier phases strictly in order (a warp that waited on a slot
                                  // shared with others could be two phases ahead of it, which the parity wait cannot tell apart)
    int gran_floats;              // T * stride
    int cap;                      // candidate buffer entries
    int compact_at;               // compact when a candidate lands at or beyond this position
    int max_segs;                 // images one CTA range can touch
    int seg_keys;                 // keys a segment can publish (>= K)
    unsigned long long* keys;     // [grid][max_segs][seg_keys]
    SegMeta* segmeta;             // [grid][max_segs]
    unsigned long long* hint;     // [2] in the workspace: {cookie, predicted threshold bits} left by the previous call (a cache:
                                  // any content is safe, a wrong prediction only costs the slow path of the merge kernel)
    unsigned long long cookie;
    int rescan_step;              // appends after which the score histogram is scanned again
    unsigned int thr0_bits;       // threshold a segment starts with (0: none, bootstrap with hints)
    int ring;                     // 1: the 3x3 neighbours of a record are read from the ring (granules stay resident until the
                                  // records of their neighbourhood are tested); 0: from global memory (maps too wide for that)
    int hg;                       // ring mode: granules of history / lookahead that hold a pixel's 3x3 neighbourhood
    float inv_T, inv_W;           // 1 / T, 1 / W
    unsigned char* seg_out;       // fused semseg argmax: class ids [B*H*W] (NULL: off)
    int seg_off, seg_n;           // channels [seg_off, seg_off + seg_n) of every pixel
};

// Fixed-size head of the dynamic shared memory block; the ring, the queues, the candidate buffer, the select scratch and
// the score histogram follow (see the sm_*() accessors).
struct SharedHead {
    uint64_t full_bar[kMaxSlots];   // loader -> scanners: granule arrived
    uint64_t empty_bar[kMaxSlots];  // owning scanner warp -> loader: granule scanned
    unsigned int hist[256];       // radix select
    int misc[4];
    unsigned int thr_bits[2];     // per segment parity: running threshold score (float bits); 0 = not established (hint phase)
    float hint_guess[2];          // pixels whose maximum reaches this were hinted (first granule of the segment)
    unsigned int thr_init[2];     // what the segment started with: 0 = hints (exact bootstrap), else a PROVISIONAL threshold
                                  // predicted from the previous segment (see flush_segment)
    unsigned int next_thr;        // flush_segment: prediction for the next segment
    unsigned int pub_thr;         // flush_segment: score bound of the published keys
    int target[2];                // peaks the running threshold keeps at or above it (K, or 2K behind a provisional threshold)
    int hint_done[2];             // tester warps that have seen the end of the segment's hints
    int flushed;                  // segments published so far
    int cur_par;                  // parity of the segment the testers are working on
    int compact_flag;             // an append landed at/after the compaction mark
    int count;                    // candidates buffered
    int scanned;                  // `count` when the histogram was last scanned
    int maxbin;                   // highest score bin seen so far
    int seg_done;                 // tester warps that finished the current segment
    unsigned long long thr;       // K-th key of the last exact select
    // ring mode, per slot (sequence numbers are 1 + the index of the granule in this CTA's load order, so 0 = never)
    int landed_seq[kMaxSlots];    // the slot's tenant has arrived (set by the owning scanner warp)
    // per granule, indexed by its load-order index & (kCtr - 1): twice the ring depth, because the records of a granule may
    // still be pending when its slot has already been handed to the granule kMaxSlots later
    int scan_seq[kCtr];           // the granule has been scanned: `pushed` is final
    int pushed[kCtr];             // records pushed for the granule
    int tested[kCtr];             // records of the granule the testers are done with
#ifdef CVM_DECODE_STATS
    unsigned long long stats[kThreads / 32][24];   // private to each warp's lane 0: plain adds
    int roles_done;
#endif
    unsigned int q_tail[kTestWarps][2];   // records reserved (producers)
    unsigned int q_head[kTestWarps][2];   // records consumed (tester), for back-pressure
};
constexpr int kHeadBytes = (sizeof(SharedHead) + 127) & ~127;
constexpr int kQueueBytes = kTestWarps * 2 * kQCap * 8;

// The dynamic shared memory block, declared at file scope so that every device function addresses it in the shared
// state space (LDS/STS/ATOMS); passing views of it through structs or pointers degrades them to generic accesses.
extern __shared__ __align__(128) unsigned char g_smem[];

__host__ __device__ __forceinline__ size_t ring_bytes_of(int S, int gran_floats) {
    return (((size_t)S * gran_floats + 4) * 4 + 15) & ~(size_t)15;
}
__device__ __forceinline__ SharedHead* sm_head() { return reinterpret_cast<SharedHead*>(g_smem); }
// [S][gran_floats] (+ 4 floats of slack for vector over-reads of the last pixel)
__device__ __forceinline__ float* sm_ring() { return reinterpret_cast<float*>(g_smem + kHeadBytes); }
// [kTestWarps][2][kQCap] record queues
__device__ __forceinline__ unsigned long long* sm_queue(const DecodeParams& p, int t, int par) {
    return reinterpret_cast<unsigned long long*>(g_smem + kHeadBytes + ring_bytes_of(p.S, p.gran_floats)) + (size_t)(t * 2 + par) * kQCap;
}
// [cap] candidate keys
__device__ __forceinline__ unsigned long long* sm_cand(const DecodeParams& p) {
    return reinterpret_cast<unsigned long long*>(g_smem + kHeadBytes + ring_bytes_of(p.S, p.gran_floats) + kQueueBytes);
}
// [K] select scratch
__device__ __forceinline__ unsigned long long* sm_keep(const DecodeParams& p) { return sm_cand(p) + p.cap; }
// [kScoreBins] score histogram of the buffered candidates
__device__ __forceinline__ unsigned int* sm_shist(const DecodeParams& p) {
    return reinterpret_cast<unsigned int*>(sm_keep(p) + p.K);
}

size_t smem_bytes(int S, int gran_floats, int cap, int K) {
    return (size_t)kHeadBytes + ring_bytes_of(S, gran_floats) + kQueueBytes + (size_t)cap * 8 + (size_t)K * 8 + (size_t)kScoreBins * 4;
}

// Ordering between the warps of the CTA.  All flags, counters and queue records live in shared memory and are accessed with
// volatile loads / stores or atomics, which the SM performs in each warp's program order; a consumer only acts on a value
// after it has read the flag that guards it (a control dependency).  fence_cta() (acquire-release at CTA scope) is kept at
// the rare hand-over points; order_only() just stops the compiler from moving accesses.  (__threadfence_block() compiles
// to MEMBAR.SC.CTA, a sequentially consistent fence that waits out everything the thread has in flight: thousands of
// cycles per call next to a saturated memory system.)
__device__ __forceinline__ void fence_cta() { asm volatile("fence.acq_rel.cta;" ::: "memory"); }
__device__ __forceinline__ void order_only() { asm volatile("" ::: "memory"); }

// ring slot of load-order index x (x may be a little below zero for offsets that are never dereferenced)
// (ring mode runs with 8 slots or 7, see plan_decode: a mask or a constant-divisor remainder.  A generic `x % S` here - it
// is evaluated for every neighbour offset of every record - cost 16 % of the kernel.)
// Only the instantiation for 64-byte pixels (CenterTracker: whole rows need the room of 7 slots) runs with a slot count that
// is not a compile-time 8.
template <int STRIDE, int HM>
__device__ __forceinline__ int ring_slots(const DecodeParams& p) { return (STRIDE == 16 && HM == 10) ? p.S : kMaxSlots; }

__device__ __forceinline__ int ring_slot(int S, int x) {
    const unsigned u = (unsigned)(x + 7 * (1 << 20));
    return S == kMaxSlots ? (x & (kMaxSlots - 1)) : (int)(u % 7u);
}

__device__ __forceinline__ unsigned ld_vol(const unsigned* p) { return *(volatile const unsigned*)p; }
__device__ __forceinline__ int ld_vol(const int* p) { return *(volatile const int*)p; }

// Every wait loop of the kernel is bounded: a protocol error traps (the launch fails loudly) instead of hanging the GPU.
#define CVM_SPIN_LIMIT (1 << 24)
#define SPIN_GUARD(spins, what)                                                                                           \
    do {                                                                                                                  \
        if (++(spins) > CVM_SPIN_LIMIT) {                                                                                 \
            if ((threadIdx.x & 31) == 0)                                                                                  \
                printf("decode_scan_kernel: stuck in %s (cta %d warp %d)\n", what, (int)blockIdx.x, (int)(threadIdx.x >> 5)); \
            __trap();                                                                                                     \
        }                                                                                                                 \
    } while (0)

__device__ __forceinline__ void mbar_wait_guarded(uint64_t* bar, uint32_t parity, const char* what) {
    int spins = 0;
    while (!mbar_try_wait(bar, parity)) SPIN_GUARD(spins, what);
}

// -DCVM_DECODE_STATS: per-warp cycle / event counters for tuning (tools/decode_stats.py); never in the shipped build
#ifdef CVM_DECODE_STATS
__device__ unsigned long long g_decode_stats[32];
__device__ unsigned long long g_decode_cta[4][256];   // per CTA: globaltimer at start, loader done, scanners done, testers done
__device__ __forceinline__ unsigned long long gtime() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
#define STAT_T0() const long long st_t0_ = clock64()
#define STAT_LAP(i) do { const long long t_ = clock64(); if ((threadIdx.x & 31) == 0) sm_head()->stats[threadIdx.x >> 5][i] += (unsigned long long)(t_ - st_lap_); st_lap_ = t_; } while (0)
#define STAT_LAP0() long long st_lap_ = clock64()
// (accumulated in shared memory, flushed to the global array by the last warp of the CTA to finish: global atomics inside the
// polling loops would disturb what is being measured)
#define STAT_ACC(i) do { if ((threadIdx.x & 31) == 0) sm_head()->stats[threadIdx.x >> 5][i] += (unsigned long long)(clock64() - st_t0_); } while (0)
#define STAT_ADD(i, v) do { if ((threadIdx.x & 31) == 0) sm_head()->stats[threadIdx.x >> 5][i] += (unsigned long long)(v); } while (0)
#else
#define STAT_T0() ((void)0)
#define STAT_LAP(i) ((void)0)
#define STAT_LAP0() ((void)0)
#define STAT_ACC(i) ((void)0)
#define STAT_ADD(i, v) ((void)0)
#endif

#ifdef CVM_DECODE_STATS
__device__ __noinline__ void stats_role_done(const DecodeParams& p) {
    __syncwarp();
    if ((threadIdx.x & 31) == 0) {
        __threadfence_block();
        if (atomicAdd(&sm_head()->roles_done, 1) == kTestWarps + p.S) {
            __threadfence_block();
            for (int w = 0; w < kThreads / 32; ++w)
                for (int k = 0; k < 24; ++k) atomicAdd(&g_decode_stats[k], sm_head()->stats[w][k]);
        }
    }
}
#endif

// named barrier over the first `nt` threads of the CTA (the tester warps; scanners and loader never join)
__device__ __forceinline__ void group_sync(int nt) { asm volatile("bar.sync 1, %0;" ::"r"(nt) : "memory"); }

__device__ __noinline__ void gather(const DecodeParams& p, bool at_segment_end, int seg);

// ---- exact top-K select on distinct 64-bit keys held in shared memory -------------------------------------------------
// On return keys[0..K) hold the K largest (unordered), *thr is the K-th largest key.  n > K required.  Called by the
// first nt threads of the CTA (nt >= 32).
__device__ __noinline__ void select_topk(unsigned long long* keys, int n, int K, unsigned int* hist, unsigned long long* keep,
                                         int* s_misc /* [4] */, unsigned long long* thr_out, int nt) {
    const int tid = threadIdx.x;
    unsigned long long prefix = 0ull, mask = 0ull;
    int need = K;
    for (int shift = 56; shift >= 0; shift -= 8) {
        for (int i = tid; i < 256; i += nt) hist[i] = 0;
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
                const unsigned int t = __shfl_up_sync(kFull, incl, o);
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
// entry is a real peak of this segment, so a later candidate below the lower edge of tb cannot make the top K: the edge
// becomes the new threshold.  Counts read while other warps append are at worst too low, which only makes the bound
// conservative.  No buffer traffic, no barrier.
__device__ __forceinline__ void scan_threshold(SharedHead* h, unsigned int* thr_word, const unsigned int* shist, int K, int lane) {
    int n = 0, top = 0;
    if (lane == 0) {   // one lane decides (the words change under our feet), the warp follows
        n = ld_vol(&h->count);
        top = ld_vol(&h->maxbin);
        if (n == ld_vol(&h->scanned) || n < K) n = -1;
    }
    n = __shfl_sync(kFull, n, 0);
    top = __shfl_sync(kFull, top, 0);
    if (n < 0) return;
    unsigned above = 0;
    for (int base = top; base >= 0; base -= 32) {
        const int bin = base - lane;
        const unsigned v = bin >= 0 ? ((volatile const unsigned int*)shist)[bin] : 0u;
        unsigned incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned t = __shfl_up_sync(kFull, incl, o);
            if (lane >= o) incl += t;
        }
        const unsigned hit = __ballot_sync(kFull, above + incl >= (unsigned)K);
        if (hit) {
            const int tb = base - (__ffs(hit) - 1);
            if (lane == 0) {
                atomicMax(thr_word, (unsigned)tb << kScoreShift);   // several warps may scan at once
                h->scanned = n;
            }
            return;
        }
        above += __shfl_sync(kFull, incl, 31);
    }
    if (lane == 0) h->scanned = n;
}

// Sum of `v` over the first nt threads of the CTA (one shared atomic per warp); `slot` is scratch.
__device__ __forceinline__ int group_sum(int v, int nt, int* slot) {
    if (threadIdx.x == 0) *slot = 0;
    group_sync(nt);
    v = __reduce_add_sync(kFull, v);
    if ((threadIdx.x & 31) == 0 && v) atomicAdd(slot, v);
    group_sync(nt);
    const int total = *(volatile int*)slot;
    group_sync(nt);
    return total;
}

// Fallback when the candidate buffer passed the mark: exact select, then the histogram is rebuilt from the K survivors.
// All tester threads, no appends in flight.
__device__ __noinline__ void compact_buffer(const DecodeParams& p) {
    const int tid = threadIdx.x, nt = kTestThreads, K = p.K;
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
        atomicMax(&h->thr_bits[h->cur_par], bits > 1u ? bits : 1u);
        h->compact_flag = 0;
    }
    group_sync(nt);
}

// Lower edge (float bits) of the highest score bin with at least `target` buffered candidates at or above it; 0 if there
// are fewer.  One warp, nobody appending.
__device__ __forceinline__ unsigned hist_edge(const unsigned int* shist, int top, int target, int lane) {
    unsigned above = 0;
    for (int base = top; base >= 0; base -= 32) {
        const int bin = base - lane;
        unsigned incl = bin >= 0 ? shist[bin] : 0u;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned t = __shfl_up_sync(kFull, incl, o);
            if (lane >= o) incl += t;
        }
        const unsigned hit = __ballot_sync(kFull, above + incl >= (unsigned)target);
        if (hit) return (unsigned)(base - (__ffs(hit) - 1)) << kScoreShift;
        above += __shfl_sync(kFull, incl, 31);
    }
    return 0u;
}

// End of a (CTA, image) segment: publish the candidates that still reach the running threshold (the merge kernel does
// the exact select; only a segment with more than kSegKeys of them selects its best K first) together with that bound
// and whether it is VERIFIED - at least K peaks of this segment reach it, or the segment was bootstrapped exactly -
// or still the provisional value the segment started with (then the merge kernel checks it against the whole image).
// Then reset the selection state and hand the next segment its provisional threshold: the score that about 2K peaks of
// THIS segment reach (images of a batch look alike: an image then yields about 2K records instead of the several thousand
// a bootstrap from nothing costs; if the prediction is too high for it, the merge kernel notices and recomputes that
// image the slow way).  The prediction serves the next segment BUT ONE of this CTA (the next one is already being
// scanned) and, from CTA 0, the next call.  All tester threads, every record of the segment consumed.
__device__ __noinline__ void flush_segment(const DecodeParams& p, int seg) {
    const int tid = threadIdx.x, nt = kTestThreads;
    SharedHead* h = sm_head();
    unsigned long long* cand = sm_cand(p);
    unsigned int* shist = sm_shist(p);
    const size_t o = (size_t)blockIdx.x * p.max_segs + seg;
    unsigned long long* out = p.keys + o * p.seg_keys;
    const int par = seg & 1;
    int n = h->count;
    unsigned thr_bits = h->thr_bits[par];
    if (thr_bits == 0u) thr_bits = 1u;
    const unsigned init = h->thr_init[par];
    // init 0: exact bootstrap from hints; 1: "every positive score"; anything else is provisional until a histogram scan or
    // an exact select (both need >= K peaks at or above the new value) has raised the running threshold beyond it
    bool verified = init <= 1u || thr_bits != init;
    // prediction for the next segment (warp 0; the histogram holds every peak appended in this segment)
    if (tid < 32) {
#ifndef CVM_PRED_MULT
#define CVM_PRED_MULT 3          /* in halves of K: the prediction is the score that 1.5 K peaks of the segment reach */
#endif
        int want = CVM_PRED_MULT * p.K / 2;
        if (want > p.compact_at - 64) want = p.compact_at - 64;
        unsigned next = n >= want ? hist_edge(shist, h->maxbin, want, tid) : 0u;
        if (next == 0u) {   // fewer than that many peaks seen: go a little below what this segment ended with
            const float f = __uint_as_float(thr_bits) * (n >= p.K ? 0.9f : 0.7f);
            next = __float_as_uint(f);
        }
        if (next < 1u || thr_bits <= 1u) next = 1u;   // (a segment that never got past "every positive score" predicts nothing)
        // what is worth publishing: with K peaks or more in the buffer, only the score bins that hold the best K of them
        // (a verified bound: K peaks of this segment reach it) - the merge kernel ranks what it is given
        const unsigned pub = n >= p.K ? hist_edge(shist, h->maxbin, p.K, tid) : 0u;
        if (tid == 0) {
            h->next_thr = next;
            h->pub_thr = pub;
        }
    }
    group_sync(nt);
    if (h->pub_thr > thr_bits) {
        thr_bits = h->pub_thr;
        verified = true;
    }
    int above = 0;
    for (int i = tid; i < n; i += nt) above += (unsigned)(cand[i] >> 32) >= thr_bits;
    above = group_sum(above, nt, &h->misc[0]);
    if (above > p.seg_keys) {   // rare: heavy ties at the threshold score
        select_topk(cand, n, p.K, h->hist, sm_keep(p), h->misc, &h->thr, nt);
        n = p.K;
        for (int i = tid; i < n; i += nt) out[i] = cand[i];
        thr_bits = (unsigned)(h->thr >> 32);   // every published key reaches the K-th one
        verified = true;
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
        p.segmeta[o] = SegMeta{n, thr_bits, verified ? 1u : 0u, 0u};
        const unsigned next = p.thr0_bits ? p.thr0_bits : h->next_thr;
        h->count = 0;
        h->thr = 0ull;
        h->hint_done[par] = 0;
        // the next segment (other parity) is already under way with the prediction of the segment before this one; this
        // parity's next tenant, segment seg + 2, starts here ...
        h->thr_bits[par] = next;
        h->thr_init[par] = next;              // ... and this is what it has to get past to verify it
        h->target[par] = next > 1u ? min(CVM_PRED_MULT * p.K / 2, p.compact_at - 64) : p.K;
        if (blockIdx.x == 0) {                // the prediction the next call starts with
            p.hint[1] = next;
            p.hint[0] = p.cookie;
        }
        h->cur_par = par ^ 1;
        h->scanned = 0;
        h->maxbin = 0;
        h->compact_flag = 0;
        h->seg_done = 0;
        fence_cta();
        *(volatile int*)&h->flushed = seg + 1;
    }
    group_sync(nt);
}

// Gathering of the tester warps (called warp-uniformly between two batches of records).  A warp comes here when it saw
// the compaction flag, or (at_segment_end) when it has consumed the segment's last record and wants the flush; whoever
// arrives waits for all of them, then everybody takes the same decisions from state that cannot change while all are paused.
__device__ __noinline__ void gather(const DecodeParams& p, bool at_segment_end, int seg) {
    SharedHead* h = sm_head();
    for (;;) {
        group_sync(kTestThreads);   // everybody paused: no appends in flight
        const int flag = ld_vol(&h->compact_flag);
        const int done = ld_vol(&h->seg_done);
        group_sync(kTestThreads);
        if (done == kTestWarps) {   // all warps finished the segment: publish it (selects when more than K are buffered)
            flush_segment(p, seg);
            return;
        }
        if (flag) compact_buffer(p);        // clears the flag
        if (!at_segment_end) return;        // back to testing; a warp waiting for the flush keeps gathering
    }
}

// channel masks: 32 bits when the compile-time channel count allows it
template <int HM>
using mask_t = typename std::conditional<(HM > 0 && HM <= 32), unsigned int, unsigned long long>::type;

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

// maximum over the HM leading floats of one pixel; STRIDE > 0: compile-time layout, widest aligned vector loads
template <int STRIDE, int HM>
__device__ __forceinline__ float pixel_max(const float* px, int hm) {
    float m = __int_as_float(0xff800000);
    if (STRIDE == 0) {
        for (int c = 0; c < hm; ++c) m = fmaxf(m, px[c]);
    } else if (STRIDE == 16 && HM == 10) {
        // 64-byte pixels: lanes 0, 2, 4, 6 of a quarter warp would hit the same banks with the same 16-byte chunk; each lane
        // pair starts at a different chunk instead (all four chunks are read: no bank conflicts, one load more)
        const int r = ((threadIdx.x & 31) >> 1) & 3;
        const float ninf = __int_as_float(0xff800000);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int cid = (r + j) & 3;
            const float4 t = *reinterpret_cast<const float4*>(px + cid * 4);
            const float half = fmaxf(t.x, t.y), full = fmaxf(half, fmaxf(t.z, t.w));
            m = fmaxf(m, cid < 2 ? full : (cid == 2 ? half : ninf));   // channels 0-7 whole chunks, 8-9 half of chunk 2
        }
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

// ---- queues ---------------------------------------------------------------------------------------------------------------
// Single consumer (tester warp t), many producers (scanner warps).  A producer reserves positions with one atomic per warp
// and push, waits until they are free (back-pressure), and writes each record with ONE 64-bit store; the lap bit of a
// record tells the consumer whether the slot holds this lap's record or is still waiting for it, so slots are never
// cleared and there is no tail to publish.
__device__ __forceinline__ void q_store(unsigned long long* q, unsigned pos, unsigned lo, unsigned hi) {
    const unsigned lap = (((pos >> kQShift) & 1u) ^ 1u) << 31;   // first lap writes 1: the zero-initialised queue reads as empty
    *(volatile unsigned long long*)(q + (pos & (kQCap - 1))) = ((unsigned long long)hi << 32) | (unsigned long long)(lo | lap);
}

__device__ __forceinline__ void q_wait_space(SharedHead* h, int t, int par, unsigned end_pos) {
    int spins = 0;
    STAT_T0();
    while ((int)(end_pos - ld_vol(&h->q_head[t][par])) > kQCap) {
        __nanosleep(32);
        SPIN_GUARD(spins, "queue back-pressure");
    }
#ifdef CVM_DECODE_STATS
    h->stats[threadIdx.x >> 5][4 + (threadIdx.x & 31 ? 16 : 0)] += (unsigned long long)(clock64() - st_t0_);   // (lane 0: pushes; others: markers)
#endif
}

// one marker record to every tester queue of the parity (lane t serves queue t)
__device__ __forceinline__ void q_push_marker(const DecodeParams& p, int par, unsigned code) {
    SharedHead* h = sm_head();
    const int lane = threadIdx.x & 31;
    __syncwarp();
    if (lane < kTestWarps) {
        const unsigned pos = atomicAdd(&h->q_tail[lane][par], 1u);
        q_wait_space(h, lane, par, pos + 1u);
        order_only();
        q_store(sm_queue(p, lane, par), pos, kRecMarker | code, 0u);
    }
    __syncwarp();
}

// Ring mode: has granule `ls` of the load order landed in its slot?  Its owner publishes that in landed_seq the moment it sees
// it; before that, anybody may ask the slot's full barrier - but only once the slot's previous tenant (ls - kMaxSlots) is
// known to have landed, because the parity test cannot tell phases two apart.
__device__ __forceinline__ bool granule_landed(SharedHead* h, int S, int ls) {
    const int s = ring_slot(S, ls);
    const int seq = ld_vol(&h->landed_seq[s]);
    if (seq >= ls + 1) return true;
    if (seq >= ls + 1 - S) {
        const int turn = S == kMaxSlots ? ls >> 3 : (int)((unsigned)ls / 7u);
        return mbar_try_wait(&h->full_bar[s], (uint32_t)(turn & 1));
    }
    return false;
}

// ---- scanner warps ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void wait_flushed(SharedHead* h, int want) {
    if (want <= 0) return;
    int spins = 0;
    STAT_T0();
    while (ld_vol(&h->flushed) < want) {
        __nanosleep(64);
        SPIN_GUARD(spins, "wait for a segment flush");
    }
    STAT_ACC(3);
}

// Scan of one granule: bit j of the result = pixel (j / kUnroll) * 32 * kUnroll + (j % kUnroll) * 32 + lane of the granule has
// a heatmap channel that reaches thr_f.  Nothing else happens per pixel: which channels, and whether they are peaks, is the
// testers' business.  ARGMAX: also write the class id of the pixel's semseg slice.
template <int STRIDE, int HM, bool ARGMAX>
__device__ __forceinline__ unsigned scan_granule(const DecodeParams& p, const float* gran, int npx, int q0, long long img, float thr_f) {
    const int lane = threadIdx.x & 31;
    const int stride = STRIDE ? STRIDE : p.stride, hm = HM ? HM : p.hm;
    unsigned bits = 0u;
    for (int base = 0, j = 0; base < npx; base += 32 * kUnroll, j += kUnroll) {
#pragma unroll
        for (int k = 0; k < kUnroll; ++k) {
            const int pix = base + k * 32 + lane;
            const bool ok = pix < npx;
            const float* px = gran + (size_t)(ok ? pix : 0) * stride;
            const float vmax = pixel_max<STRIDE, HM>(px, hm);
            bits |= (ok && vmax >= thr_f ? 1u : 0u) << (j + k);
            if (ARGMAX && ok) {   // first maximum of the semseg slice (np.argmax, common/utils/image.py:88)
                const float* sp = px + p.seg_off;
                int idx = 0;
                float best = sp[0];
                for (int c = 1; c < p.seg_n; ++c) {
                    const float v = sp[c];
                    if (v > best) {
                        best = v;
                        idx = c;
                    }
                }
                p.seg_out[(size_t)img * p.HW + (size_t)(q0 + pix)] = (unsigned char)idx;
            }
        }
    }
    return bits;
}

// Push one record per hit bit (see scan_granule) to the tester queues: chunks of kUnroll bits per lane (<= 32 * kUnroll
// records, one reservation each).  spread: every chunk to the next tester (hints: many records, and the segment waits for
// them); otherwise all to tester `rr` (a tester finds the few records of a granule together).  Returns the records pushed.
__device__ __forceinline__ int push_hits(const DecodeParams& p, unsigned bits, int npx, int q0, int par, unsigned kind, bool spread,
                                         unsigned& rr) {
    SharedHead* h = sm_head();
    const int lane = threadIdx.x & 31;
    int n_pushed = 0;
    if (!__any_sync(kFull, bits != 0u)) return 0;
    for (int base = 0, j = 0; base < npx; base += 32 * kUnroll, j += kUnroll) {
        const unsigned mine = (bits >> j) & ((1u << kUnroll) - 1u);
        if (!__any_sync(kFull, mine != 0u)) continue;
        // exclusive prefix of the per-lane counts
        const int cnt = __popc(mine);
        int incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(kFull, incl, o);
            if (lane >= o) incl += v;
        }
        const int total = __shfl_sync(kFull, incl, 31);
        const int t = (int)(rr % kTestWarps);
        if (spread) ++rr;
        unsigned pos = 0;
        if (lane == 0) {
            pos = atomicAdd(&h->q_tail[t][par], (unsigned)total);
            q_wait_space(h, t, par, pos + (unsigned)total);
        }
        pos = __shfl_sync(kFull, pos, 0) + (unsigned)(incl - cnt);
        order_only();   // (the consumer's reads of the old tenants happened before its head update that lane 0 saw)
        unsigned long long* q = sm_queue(p, t, par);
#pragma unroll
        for (int k = 0; k < kUnroll; ++k) {
            if (mine & (1u << k)) q_store(q, pos++, (unsigned)(q0 + base + k * 32 + lane), kind);
        }
        n_pushed += total;
    }
    return n_pushed;
}

// Ring mode.  For every hit bit (see scan_granule) of the granule in `slot`: pick the pixel's channels (kKindAll: those
// that reach the threshold; kKindHint: its best one; kKindRest: like kKindAll without the one a hint covered) and test each
// against the five neighbours that are already in the ring - the row above and the pixel's left / right neighbours (the
// granules before this one are still resident, see the loader).  What survives is pushed to the testers as (pixel, channel,
// score) and only waits for the row below.  Lanes walk their own hit pixels; every round each lane tests one (pixel,
// channel) pair and the warp pushes the survivors with one reservation.  Returns the records pushed.
template <int STRIDE, int HM>
__device__ __forceinline__ int prefilter_push(const DecodeParams& p, int slot, int npx, int q0, unsigned bits, int par, unsigned kind,
                                              float guess, bool spread, unsigned& rr) {
    using M = mask_t<HM>;
    SharedHead* h = sm_head();
    const float* const ring = sm_ring();
    const int lane = threadIdx.x & 31;
    const int stride = STRIDE ? STRIDE : p.stride, hm = HM ? HM : p.hm, W = p.W, T = p.T, G = p.gran_floats;
    const float ninf = __int_as_float(0xff800000);
    int n_pushed = 0;
    M chans = 0;                       // channels of the lane's current pixel that are still to test
    int q = 0, own = 0, o_ul = 0, o_um = 0, o_ur = 0, o_l = 0, o_r = 0;
    bool up = false, lf = false, rt = false, rt_here = false;
    float thr_f = 0.f;
    while (__any_sync(kFull, chans != 0 || bits != 0u)) {
        if (chans == 0 && bits != 0u) {
            const int j = __ffs((int)bits) - 1;
            bits &= bits - 1u;
            const int pix = (j / kUnroll) * (32 * kUnroll) + (j % kUnroll) * 32 + lane;
            q = q0 + pix;
            int y = __float2int_rz(__fmul_rn((float)q, p.inv_W)), x = q - y * W;   // exact after one correction step
            if (x < 0) {
                --y;
                x += W;
            } else if (x >= W) {
                ++y;
                x -= W;
            }
            up = y > 0;
            lf = x > 0;
            rt = x < W - 1;
            rt_here = rt && pix < T - 1;   // (the right neighbour of a granule's last pixel is in the NEXT granule: the tester's)
            own = slot * G + pix * stride;
            o_l = pix > 0 ? own - stride : ring_slot(ring_slots<STRIDE, HM>(p), slot - 1) * G + (T - 1) * stride;
            o_r = own + stride;
            int o_up = pix - W, s_up = slot;
            while (o_up < 0) {
                o_up += T;
                s_up = ring_slot(ring_slots<STRIDE, HM>(p), s_up - 1);
            }
            o_um = s_up * G + o_up * stride;
            o_ul = o_up > 0 ? o_um - stride : ring_slot(ring_slots<STRIDE, HM>(p), s_up - 1) * G + (T - 1) * stride;
            o_ur = o_up < T - 1 ? o_um + stride : ring_slot(ring_slots<STRIDE, HM>(p), s_up + 1) * G;
            unsigned tb = ld_vol(&h->thr_bits[par]);
            tb = tb ? tb : 1u;
            thr_f = __uint_as_float(tb);
            const float* px = ring + own;
            float vmax = ninf;
            M ge = 0, top = 0;
            if (HM > 0) {
                float c[HM > 0 ? HM : 1];
#pragma unroll
                for (int k = 0; k < HM; ++k) c[k] = px[k];
#pragma unroll
                for (int k = 0; k < HM; ++k) vmax = fmaxf(vmax, c[k]);
#pragma unroll
                for (int k = 0; k < HM; ++k) {
                    ge |= (M)(c[k] >= thr_f ? 1u : 0u) << k;
                    top |= (M)(c[k] >= vmax ? 1u : 0u) << k;
                }
            } else {
                for (int k = 0; k < hm; ++k) vmax = fmaxf(vmax, px[k]);
                for (int k = 0; k < hm; ++k) {
                    ge |= (M)(px[k] >= thr_f ? 1u : 0u) << k;
                    top |= (M)(px[k] >= vmax ? 1u : 0u) << k;
                }
            }
            top &= (M)0 - top;   // lowest channel holding the maximum
            chans = kind == kKindHint ? top : ge;
            if (kind == kKindRest && vmax >= guess) chans &= ~top;
        }
        bool cand = false;
        int ch = 0;
        float v = 0.f;
        if (chans != 0) {
            ch = sizeof(M) == 4 ? __ffs((int)chans) - 1 : __ffsll((long long)chans) - 1;
            chans &= chans - 1;
            v = ring[own + ch];
            const float a0 = (up && lf) ? ring[o_ul + ch] : ninf, a1 = up ? ring[o_um + ch] : ninf;
            const float a2 = (up && rt) ? ring[o_ur + ch] : ninf;
            const float a3 = lf ? ring[o_l + ch] : ninf, a4 = rt_here ? ring[o_r + ch] : ninf;
            const float m5 = fmaxf(fmaxf(fmaxf(a0, a1), fmaxf(a2, a3)), a4);
            cand = v >= thr_f && m5 <= v;
        }
        const unsigned bal = __ballot_sync(kFull, cand);
        if (bal) {
            const int n = __popc(bal);
            const int t = (int)(rr % kTestWarps);
            if (spread) ++rr;
            unsigned pos = 0;
            if (lane == 0) {
                pos = atomicAdd(&h->q_tail[t][par], (unsigned)n);
                q_wait_space(h, t, par, pos + (unsigned)n);
            }
            pos = __shfl_sync(kFull, pos, 0);
            order_only();
            if (cand) q_store(sm_queue(p, t, par), pos + __popc(bal & ((1u << lane) - 1u)), (unsigned)q | ((unsigned)ch << 24), __float_as_uint(v));
            n_pushed += n;
        }
    }
    return n_pushed;
}

template <int STRIDE, int HM, bool SEG>
__device__ __forceinline__ void scanner_main(const DecodeParams& p, int w, long long g_first, int n_local, int lead, int n_load,
                                             long long img0, int n_segs) {
    SharedHead* const h = sm_head();
    const float* const ring = sm_ring();
    const int lane = threadIdx.x & 31;
    const int stride = STRIDE ? STRIDE : p.stride, hm = HM ? HM : p.hm;
    const float ninf = __int_as_float(0xff800000);
#ifdef CVM_DECODE_STATS
    const long long sc_t0 = clock64();
#endif
    const int nsw = p.S, slot = w;   // granule ls of the load order lives in slot ls % S == w: this warp's own slot
    int cur_seg = 0;              // segments [0, cur_seg) have got this warp's END marker
    unsigned rr = (unsigned)w;    // round-robin over the tester queues
    // load-order index ls <-> granule c = ls - lead of the CTA's range (c < 0: history before the range, c >= n_local:
    // lookahead after it; both only in ring mode, both inside the first / last image of the range)
    long long img = (g_first - lead + w) / p.gpi;
    int gi = (int)((g_first - lead + w) - img * p.gpi);
    for (int ls = w; ls < n_load; ls += nsw) {
        const int c = ls - lead;
        {
            STAT_T0();
            mbar_wait_guarded(&h->full_bar[slot], (uint32_t)((ls / nsw) & 1), "wait for a granule");
            STAT_ACC(0);
        }
        if (p.ring && lane == 0) *(volatile int*)&h->landed_seq[slot] = ls + 1;
        if (c >= 0 && c < n_local) {
            const int seg = (int)(img - img0), par = seg & 1;
            for (; cur_seg < seg; ++cur_seg) {
                wait_flushed(h, cur_seg - 1);                 // queue parity of segment cur_seg - 2 is free again
                q_push_marker(p, cur_seg & 1, kMarkEnd);
            }
            wait_flushed(h, seg - 1);    // (the flush of segment seg - 2 frees this parity and leaves its provisional threshold)
            const int q0 = gi * p.T, npx = min(p.T, p.HW - q0);
            const float* gran = ring + (size_t)slot * p.gran_floats;
            const bool first = gi == 0 || c == 0;          // first granule of the segment: its owner bootstraps the threshold
            const bool hinting = first && ld_vol(&h->thr_bits[par]) == 0u;
            unsigned bits = 0u;
            float guess = ninf;
            if (hinting) {
                // hints: the best channel of the better half of the pixels (16th largest of the first 32 pixel maxima; every
                // pixel when K is large against the granule), enough to find K high peaks fast
                if (p.K * 4 <= npx) {
                    const float v = lane < npx ? pixel_max<STRIDE, HM>(gran + (size_t)lane * stride, hm) : ninf;
                    int rank = 0;
#pragma unroll
                    for (int j = 0; j < 32; ++j) rank += __shfl_sync(kFull, v, j) > v;
                    float cc = rank >= 15 ? v : ninf;
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) cc = fmaxf(cc, __shfl_xor_sync(kFull, cc, o));
                    guess = cc;
                }
                if (lane == 0) *(volatile float*)&h->hint_guess[par] = guess;
                bits = scan_granule<STRIDE, HM, SEG>(p, gran, npx, q0, img, guess);
            } else {
                int spins = 0;
                STAT_T0();
                while (ld_vol(&h->thr_bits[par]) == 0u) {
                    __nanosleep(64);
                    SPIN_GUARD(spins, "wait for the segment's threshold");
                }
                STAT_ACC(2);
                STAT_LAP0();
                bits = scan_granule<STRIDE, HM, SEG>(p, gran, npx, q0, img, __uint_as_float(ld_vol(&h->thr_bits[par])));
                STAT_LAP(20);
            }
            if (p.ring) {
                // a record waits for the row below only in the testers' hands: push (and, before that, run the tests against
                // the rows above) only once the granules that hold the pixels after this one's (same image) have landed too.
                // Nothing here waits for a tester, and the testers never wait for a granule: no cycle.
                for (int k = 1; k <= p.hg && gi + k < p.gpi; ++k) {
                    int spins = 0;
                    STAT_T0();
                    while (!__all_sync(kFull, granule_landed(h, nsw, ls + k))) {
                        __nanosleep(32);
                        SPIN_GUARD(spins, "wait for the lookahead granule");
                    }
                    STAT_ACC(16);
                }
            }
            int n_pushed = 0;
            if (hinting) {
                n_pushed += p.ring ? prefilter_push<STRIDE, HM>(p, slot, npx, q0, bits, par, kKindHint, guess, true, rr)
                                   : push_hits(p, bits, npx, q0, par, kKindHint, true, rr);
                q_push_marker(p, par, kMarkHintEnd);
                int spins = 0;
                STAT_T0();
                while (ld_vol(&h->thr_bits[par]) == 0u) {
                    __nanosleep(64);
                    SPIN_GUARD(spins, "wait for the first threshold");
                }
                STAT_ACC(1);
                // second pass: everything that reaches the threshold (the testers leave out the channels the hints covered)
                bits = scan_granule<STRIDE, HM, false>(p, gran, npx, q0, img, __uint_as_float(ld_vol(&h->thr_bits[par])));
                n_pushed += p.ring ? prefilter_push<STRIDE, HM>(p, slot, npx, q0, bits, par, kKindRest, guess, true, rr)
                                   : push_hits(p, bits, npx, q0, par, kKindRest, true, rr);
            } else {
                STAT_LAP0();
                n_pushed += p.ring ? prefilter_push<STRIDE, HM>(p, slot, npx, q0, bits, par, kKindAll, guess, false, rr)
                                   : push_hits(p, bits, npx, q0, par, kKindAll, false, rr);
                STAT_LAP(21);
                ++rr;
            }
            __syncwarp();   // every lane is done reading the granule
            if (lane == 0) {
                if (p.ring) {   // the granule stays while the scans / records around it need it (the loader warp watches the counters)
                    *(volatile int*)&h->pushed[ls & (kCtr - 1)] = n_pushed;
                    fence_cta();
                    *(volatile int*)&h->scan_seq[ls & (kCtr - 1)] = ls + 1;
                } else {
                    mbar_arrive(&h->empty_bar[slot]);
                }
            }
        } else if (lane == 0) {   // history / lookahead granule (ring mode only): resident for its neighbours' tests, not scanned
            *(volatile int*)&h->pushed[ls & (kCtr - 1)] = 0;
            fence_cta();
            *(volatile int*)&h->scan_seq[ls & (kCtr - 1)] = ls + 1;
        }
        gi += nsw;
        while (gi >= p.gpi) {
            gi -= p.gpi;
            ++img;
        }
    }
    for (; cur_seg < n_segs; ++cur_seg) {
        wait_flushed(h, cur_seg - 1);
        q_push_marker(p, cur_seg & 1, kMarkEnd);
    }
#ifdef CVM_DECODE_STATS
    STAT_ADD(5, clock64() - sc_t0);
#endif
}

// ---- tester warps -------------------------------------------------------------------------------------------------------
// Append the peaks of one round (warp-uniform call; pm = ballot of `peak`, non-zero) to the candidate buffer: one shared
// atomic per warp, keys placed by ballot rank.  Ends at a safe point: joins a pending compaction, or rescans the score
// histogram when enough keys have come in.
__device__ __forceinline__ void append_peaks(const DecodeParams& p, int par, bool peak, float v, unsigned flat, unsigned pm) {
    SharedHead* h = sm_head();
    unsigned long long* cand = sm_cand(p);
    unsigned int* shist = sm_shist(p);
    const int lane = threadIdx.x & 31;
    const unsigned total = __popc(pm);
    unsigned base = 0;
    int rescan = 0;
    if (lane == 0) {
        base = (unsigned)atomicAdd(&h->count, (int)total);
        const int cnt = (int)(base + total);
        rescan = cnt >= ld_vol(&h->target[par]) && cnt - ld_vol(&h->scanned) >= p.rescan_step;
        if (cnt > p.cap) {   // cannot happen: every warp stops within one round of the mark (see plan_decode)
            printf("decode_scan_kernel: candidate buffer overflow (cta %d)\n", (int)blockIdx.x);
            __trap();
        }
    }
    base = __shfl_sync(kFull, base, 0);
    rescan = __shfl_sync(kFull, rescan, 0);
    if (peak) {
        const unsigned pos = base + __popc(pm & ((1u << lane) - 1u));
        const unsigned bits = __float_as_uint(v);
        cand[pos] = ((unsigned long long)bits << 32) | (unsigned long long)(0xFFFFFFFFu - flat);
        if (pos >= (unsigned)p.compact_at) *(volatile int*)&h->compact_flag = 1;
        const unsigned bin = bits >> kScoreShift;
        atomicAdd(&shist[bin], 1u);
        atomicMax(&h->maxbin, (int)bin);
    }
    __syncwarp();
    if (__any_sync(kFull, ld_vol(&h->compact_flag) != 0)) gather(p, false, 0);
    else if (rescan) scan_threshold(h, &h->thr_bits[par], shist, ld_vol(&h->target[par]), lane);
}

// One round of the 3x3 test: lanes with `act` test channel `ch` (relative to the record's channel base) of their record.
// RING: the nine values come from the shared-memory ring, otherwise from global memory (L2).
template <bool RING>
__device__ __forceinline__ void load_window(const float* src, const int (&off)[9], const bool (&ok)[9], int ch, float& v, float& m) {
    const float ninf = __int_as_float(0xff800000);
    float a[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) {
        if (RING) a[k] = ok[k] ? src[off[k] + ch] : ninf;
        else a[k] = ok[k] ? __ldcg(src + off[k] + ch) : ninf;
    }
    v = a[4];
    m = fmaxf(fmaxf(fmaxf(a[0], a[1]), fmaxf(a[2], a[3])), fmaxf(fmaxf(a[5], a[6]), fmaxf(a[7], a[8])));
}

// Exact 3x3 test of one batch of records (one record per lane; lanes without one pass has = false).  A record names a
// pixel whose best channel reached the threshold when it was scanned.  The lane reads the pixel's heatmap channels, keeps
// the ones that (still) reach the threshold, and every round each record tests ONE of its pending channels: nine
// independent loads per lane - from the ring (the granules around a record stay resident until it is tested:
// shared-memory latency) or, for maps too wide for that, from global memory (L2 hits) - then a ballot and one aggregated
// append per warp.  seg_ls0: load-order index of granule 0 of the records' image.
template <int STRIDE, int HM>
__device__ __forceinline__ void test_records(const DecodeParams& p, const float* img_base, int seg_ls0, int par, bool has, unsigned lo,
                                             unsigned kind) {
    using M = mask_t<HM>;
    if (!__any_sync(kFull, has)) return;
    STAT_LAP0();
    SharedHead* h = sm_head();
    const int stride = STRIDE ? STRIDE : p.stride, hm = HM ? HM : p.hm, W = p.W, T = p.T;
    const int q = has ? (int)(lo & kRecPixel) : 0;
    // row / column and (ring mode) granule / position without integer divisions (exact after one correction step)
    int y = __float2int_rz(__fmul_rn((float)q, p.inv_W)), x = q - y * W;
    if (x < 0) {
        --y;
        x += W;
    } else if (x >= W) {
        ++y;
        x -= W;
    }
    const bool up = y > 0, dn = y < p.H - 1, lf = x > 0, rt = x < W - 1;
    const bool ok[9] = {up && lf, up, up && rt, lf, true, rt, dn && lf, dn, dn && rt};
    // element offsets of the nine pixels (channel 0): into the ring, or relative to the image in global memory
    int off[9], g_own = 0;
    if (p.ring) {
        int g = __float2int_rz(__fmul_rn((float)q, p.inv_T)), o = q - g * T;
        if (o < 0) {
            --g;
            o += T;
        } else if (o >= T) {
            ++g;
            o -= T;
        }
        g_own = g;
        int gr[3] = {g, g, g}, orow[3] = {o - W, o, o + W};   // granule / position of the pixels above, at and below
        while (orow[0] < 0) {
            orow[0] += T;
            --gr[0];
        }
        while (orow[2] >= T) {
            orow[2] -= T;
            ++gr[2];
        }
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            const int base = ring_slot(ring_slots<STRIDE, HM>(p), seg_ls0 + gr[r]) * p.gran_floats;
            off[r * 3 + 1] = base + orow[r] * stride;
            // left / right neighbour: the same granule unless the pixel sits at its edge
            off[r * 3 + 0] = orow[r] > 0 ? off[r * 3 + 1] - stride
                                         : ring_slot(ring_slots<STRIDE, HM>(p), seg_ls0 + gr[r] - 1) * p.gran_floats + (T - 1) * stride;
            off[r * 3 + 2] = orow[r] < T - 1 ? off[r * 3 + 1] + stride : ring_slot(ring_slots<STRIDE, HM>(p), seg_ls0 + gr[r] + 1) * p.gran_floats;
        }
    } else {
#pragma unroll
        for (int k = 0; k < 9; ++k) off[k] = (q + (k / 3 - 1) * W + (k % 3 - 1)) * stride;
    }
    const float* const ring = sm_ring();
    // the channels of the pixel to test
    unsigned tb = ld_vol(&h->thr_bits[par]);
    tb = tb ? tb : 1u;
    M rem = 0;
    if (has) {
        const float thr_f = __uint_as_float(tb);
        const float* own = p.ring ? ring + off[4] : img_base + off[4];
        float vmax = __int_as_float(0xff800000);
        M ge = 0;
        if (HM > 0) {
            float c[HM > 0 ? HM : 1];
#pragma unroll
            for (int k = 0; k < HM; ++k) c[k] = p.ring ? own[k] : __ldcg(own + k);
#pragma unroll
            for (int k = 0; k < HM; ++k) vmax = fmaxf(vmax, c[k]);
            M top = 0;
#pragma unroll
            for (int k = 0; k < HM; ++k) {
                ge |= (M)(c[k] >= thr_f ? 1u : 0u) << k;
                top |= (M)(c[k] >= vmax ? 1u : 0u) << k;
            }
            top &= (M)0 - top;   // lowest channel holding the maximum
            rem = kind == kKindHint ? top : ge;
            if (kind == kKindRest && vmax >= *(volatile const float*)&h->hint_guess[par]) rem &= ~top;
        } else {
            M top = 0;
            for (int k = 0; k < hm; ++k) vmax = fmaxf(vmax, p.ring ? own[k] : __ldcg(own + k));
            for (int k = 0; k < hm; ++k) {
                const float v = p.ring ? own[k] : __ldcg(own + k);
                ge |= (M)(v >= thr_f ? 1u : 0u) << k;
                top |= (M)(v >= vmax ? 1u : 0u) << k;
            }
            top &= (M)0 - top;
            rem = kind == kKindHint ? top : ge;
            if (kind == kKindRest && vmax >= *(volatile const float*)&h->hint_guess[par]) rem &= ~top;
        }
    }
    STAT_LAP(17);
    while (__any_sync(kFull, rem != 0)) {
        const bool act = rem != 0;
        int ch = 0;
        float v = 0.f, m = __int_as_float(0xff800000);
        if (act) {
            ch = sizeof(M) == 4 ? __ffs((int)rem) - 1 : __ffsll((long long)rem) - 1;
            rem &= rem - 1;
            if (p.ring) load_window<true>(ring, off, ok, ch, v, m);
            else load_window<false>(img_base, off, ok, ch, v, m);
        }
        // the threshold may have moved since the scan: only scores that still reach it (and are > 0) go in
        tb = ld_vol(&h->thr_bits[par]);
        tb = tb ? tb : 1u;
        const bool peak = act && v >= __uint_as_float(tb) && m <= v;
        const unsigned pm = __ballot_sync(kFull, peak);
        STAT_LAP(18);
        STAT_ADD(11, 1);
        STAT_ADD(12, __popc(pm));
        if (pm) append_peaks(p, par, peak, v, (unsigned)q * (unsigned)hm + (unsigned)ch, pm);
        STAT_LAP(19);
    }
    // ring mode: this record no longer needs its neighbourhood (the loader warp releases a granule when the records of the
    // granules around it are all accounted for)
    if (p.ring && has) atomicAdd(&h->tested[(seg_ls0 + g_own) & (kCtr - 1)], 1);
}

// Ring mode: a record is a (pixel, channel, score) that already beat the row above and its left / right neighbours
// (prefilter_push).  What is left is the row below (and the right neighbour when that sits in the next granule): wait until
// the granules that hold them have landed (their owners publish that the moment they see them), three or four shared loads,
// a ballot and one aggregated append per warp.
template <int STRIDE, int HM>
__device__ __forceinline__ void confirm_records(const DecodeParams& p, int seg_ls0, int par, bool has, unsigned lo, unsigned hi) {
    if (!__any_sync(kFull, has)) return;
    STAT_LAP0();
    SharedHead* h = sm_head();
    const float* const ring = sm_ring();
    const int stride = STRIDE ? STRIDE : p.stride, hm = HM ? HM : p.hm, W = p.W, T = p.T, G = p.gran_floats;
    const float ninf = __int_as_float(0xff800000);
    const int q = has ? (int)(lo & 0xFFFFFFu) : 0, ch = (int)((lo >> 24) & 63u);
    const float v = __uint_as_float(hi);
    int y = __float2int_rz(__fmul_rn((float)q, p.inv_W)), x = q - y * W;   // exact after one correction step
    if (x < 0) {
        --y;
        x += W;
    } else if (x >= W) {
        ++y;
        x -= W;
    }
    int g = __float2int_rz(__fmul_rn((float)q, p.inv_T)), o = q - g * T;
    if (o < 0) {
        --g;
        o += T;
    } else if (o >= T) {
        ++g;
        o -= T;
    }
    const bool dn = has && y < p.H - 1, lf = x > 0, rt = x < W - 1;
    const bool right_later = has && rt && o == T - 1;    // the right neighbour sits in the next granule
    int o_dn = o + W, g_dn = g;
    while (o_dn >= T) {
        o_dn -= T;
        ++g_dn;
    }
    // (the scanner pushed this record only after the granules of the row below had landed: nothing to wait for)
    STAT_LAP(17);
    float m = ninf;
    if (dn) {
        const int o_dm = ring_slot(ring_slots<STRIDE, HM>(p), seg_ls0 + g_dn) * G + o_dn * stride + ch;
        const int o_dl = o_dn > 0 ? o_dm - stride : ring_slot(ring_slots<STRIDE, HM>(p), seg_ls0 + g_dn - 1) * G + (T - 1) * stride + ch;
        const int o_dr = o_dn < T - 1 ? o_dm + stride : ring_slot(ring_slots<STRIDE, HM>(p), seg_ls0 + g_dn + 1) * G + ch;
        const float a0 = lf ? ring[o_dl] : ninf, a1 = ring[o_dm], a2 = rt ? ring[o_dr] : ninf;
        m = fmaxf(a0, fmaxf(a1, a2));
    }
    if (right_later) m = fmaxf(m, ring[ring_slot(ring_slots<STRIDE, HM>(p), seg_ls0 + g + 1) * G + ch]);
    // the threshold may have moved since the scan: only scores that still reach it (and are > 0) go in
    unsigned tb = ld_vol(&h->thr_bits[par]);
    tb = tb ? tb : 1u;
    const bool peak = has && v >= __uint_as_float(tb) && m <= v;
    const unsigned pm = __ballot_sync(kFull, peak);
    STAT_LAP(18);
    STAT_ADD(11, 1);
    STAT_ADD(12, __popc(pm));
    if (pm) append_peaks(p, par, peak, v, (unsigned)q * (unsigned)hm + (unsigned)ch, pm);
    STAT_LAP(19);
    // this record no longer needs the ring (the loader warp releases a granule when the scans and records that read it are done)
    if (has) atomicAdd(&h->tested[(seg_ls0 + g) & (kCtr - 1)], 1);
}

template <int STRIDE, int HM>
__device__ __forceinline__ void tester_main(const DecodeParams& p, long long g_first, int lead, long long img0, int n_segs) {
    SharedHead* const h = sm_head();
    const int lane = threadIdx.x & 31, t = threadIdx.x >> 5;
    const int stride = STRIDE ? STRIDE : p.stride;
    unsigned head0 = 0, head1 = 0;   // consumed records per parity queue
#ifdef CVM_DECODE_STATS
    const long long te_t0 = clock64();
#endif
    for (int seg = 0; seg < n_segs; ++seg) {
        const int par = seg & 1;
        const unsigned long long* q = sm_queue(p, t, par);
        const float* img_base = p.yp + (size_t)(img0 + seg) * p.HW * stride;
        const int seg_ls0 = (int)((img0 + seg) * p.gpi - g_first) + lead;   // (multiples of the slot count away when negative: see plan)
        unsigned hd = par ? head1 : head0;
        int ends = 0, spins = 0;
        while (ends < p.S) {
            const unsigned pos = hd + (unsigned)lane;
            const unsigned long long r = *(volatile const unsigned long long*)(q + (pos & (kQCap - 1)));
            const unsigned lo = (unsigned)r, hi = (unsigned)(r >> 32);
            const bool valid = (lo >> 31) == (((pos >> kQShift) & 1u) ^ 1u);
            const unsigned vm = __ballot_sync(kFull, valid);
            const int n = vm == kFull ? 32 : __ffs((int)~vm) - 1;   // records are consumed in order: the valid prefix
            if (n == 0) {
                if (__any_sync(kFull, ld_vol(&h->compact_flag) != 0)) {
                    gather(p, false, 0);
                } else {
                    STAT_T0();
                    __nanosleep(CVM_POLL_NS);
                    SPIN_GUARD(spins, "tester waiting for records");
                    STAT_ACC(6);
                }
                continue;
            }
            spins = 0;
            STAT_ADD(9, n);
            STAT_ADD(10, 1);
            const bool mine = lane < n;
            const bool marker = mine && (lo & kRecMarker) != 0u;
            {
                STAT_T0();
                if (p.ring) confirm_records<STRIDE, HM>(p, seg_ls0, par, mine && !marker, lo, hi);
                else test_records<STRIDE, HM>(p, img_base, seg_ls0, par, mine && !marker, lo, hi);
                STAT_ACC(7);
            }
            ends += __popc(__ballot_sync(kFull, marker && (lo & 0xFFu) == kMarkEnd));
            if (__any_sync(kFull, marker && (lo & 0xFFu) == kMarkHintEnd)) {
                // every hint of this queue has been tested; the last tester warp to get here sets the first threshold: the
                // histogram's K-th best bin, or the smallest positive float when fewer than K peaks have turned up
                int old = 0;
                __syncwarp();
                if (lane == 0) {
                    fence_cta();
                    old = atomicAdd(&h->hint_done[par], 1);
                }
                old = __shfl_sync(kFull, old, 0);
                if (old == kTestWarps - 1) {
                    scan_threshold(h, &h->thr_bits[par], sm_shist(p), p.K, lane);
                    __syncwarp();
                    if (lane == 0) atomicMax(&h->thr_bits[par], 1u);
                }
            }
            hd += (unsigned)n;
            __syncwarp();
            order_only();
            if (lane == 0) *(volatile unsigned*)&h->q_head[t][par] = hd;
            if (__any_sync(kFull, ld_vol(&h->compact_flag) != 0)) gather(p, false, 0);
        }
        if (par) head1 = hd;
        else head0 = hd;
        __syncwarp();
        if (lane == 0) atomicAdd(&h->seg_done, 1);
        {
            STAT_T0();
            gather(p, true, seg);
            STAT_ACC(15);
        }
    }
#ifdef CVM_DECODE_STATS
    STAT_ADD(8, clock64() - te_t0);
#endif
}

// STRIDE/HM = 0: runtime pixel stride / heatmap channel count.  SEG: semseg argmax fused.  BULK: granules arrive by bulk
// async copy (needs 16-byte aligned granules); otherwise the loader warp copies them with plain loads.
template <int STRIDE, int HM, bool SEG, bool BULK>
__global__ void __launch_bounds__(kThreads, 1) decode_scan_kernel(const __grid_constant__ DecodeParams p) {
    SharedHead* const h = sm_head();
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int stride = STRIDE ? STRIDE : p.stride;

    // this CTA's contiguous range of the flat granule list
    const long long G = gridDim.x, g = blockIdx.x;
    const long long g0 = g * p.n_gran / G, g1 = (g + 1) * p.n_gran / G;
    if (g0 >= g1) return;
    const long long img0 = g0 / p.gpi, imgL = (g1 - 1) / p.gpi;
    const int n_local = (int)(g1 - g0), n_segs = (int)(imgL - img0) + 1;
    // ring mode: hg granules of history before the range and of lookahead after it (same image only) are loaded as well, so
    // that the first / last records of the range find their neighbourhood in the ring
    const int gi0 = (int)(g0 - img0 * p.gpi), giL = (int)(g1 - imgL * p.gpi);   // giL: first granule after the range in its image
    const int lead = p.ring ? min(p.hg, gi0) : 0, tail = p.ring ? min(p.hg, p.gpi - giL) : 0;
    const int n_load = lead + n_local + tail;

    if (tid == 0) {
        for (int k = 0; k < p.S; ++k) {
            mbar_init(&h->full_bar[k], 1);
            mbar_init(&h->empty_bar[k], 1);
            h->landed_seq[k] = 0;
        }
        for (int k = 0; k < kCtr; ++k) h->scan_seq[k] = h->pushed[k] = h->tested[k] = 0;
        mbar_fence_init();
        // a prediction left in the workspace by the previous call on (presumably) similar data; none: bootstrap from hints
        unsigned start = p.thr0_bits;
        if (!start && p.hint[0] == p.cookie) start = (unsigned)p.hint[1];
        if (start >= 0x7f800000u) start = 0u;   // (not a positive finite score)
        h->thr_bits[0] = h->thr_bits[1] = start;
        h->thr_init[0] = h->thr_init[1] = start;
        h->target[0] = h->target[1] = start > 1u ? min(CVM_PRED_MULT * p.K / 2, p.compact_at - 64) : p.K;
        h->hint_done[0] = h->hint_done[1] = 0;
        h->flushed = 0;
        h->cur_par = 0;
        h->compact_flag = 0;
        h->count = 0;
        h->scanned = 0;
        h->maxbin = 0;
        h->seg_done = 0;
        h->thr = 0ull;
        for (int k = 0; k < kTestWarps; ++k) h->q_tail[k][0] = h->q_tail[k][1] = h->q_head[k][0] = h->q_head[k][1] = 0u;
#ifdef CVM_DECODE_STATS
        for (int k = 0; k < kThreads / 32 * 24; ++k) (&h->stats[0][0])[k] = 0ull;
        h->roles_done = 0;
#endif
    }
    for (int k = tid; k < kScoreBins; k += kThreads) sm_shist(p)[k] = 0u;
    for (int k = tid; k < kTestWarps * 2 * kQCap; k += kThreads) sm_queue(p, 0, 0)[k] = 0ull;
    __syncthreads();   // the only CTA-wide barrier: from here on the three kinds of warps run on their own
#ifdef CVM_DECODE_STATS
    if (tid == 0) {
        g_decode_cta[0][blockIdx.x] = gtime();
        g_decode_cta[2][blockIdx.x] = g_decode_cta[3][blockIdx.x] = 0ull;
    }
#endif

    if (warp == kLoaderWarp) {
        // ---- loader warp: granule `ls` of the load order goes to slot ls % S once its previous tenant is free ----
        float* const ring = sm_ring();
        int slot = 0;
        long long l_img = img0;
        int gi = gi0 - lead;
        uint32_t e_parity = 0;   // (L2 mode) parity of the empty-barrier phase that frees a slot for its next tenant
        // ring mode: a slot is free when the records of every granule within hg of its tenant (same image) are tested.
        // tested_upto: all granules [0, tested_upto] of the load order are scanned and their records tested;
        // released: granules [0, released) may be overwritten; rel_end: first load-order index after `released`'s image
        int scan_upto = -1, tested_upto = -1, released = 0, rel_end = min(n_load, p.gpi - (gi0 - lead));
#ifdef CVM_DECODE_STATS
        const long long lo_t0 = clock64();
#endif
        for (int ls = 0; ls < n_load; ++ls) {
            if (ls >= p.S) {
                STAT_T0();
                if (!p.ring) {
                    mbar_wait_guarded(&h->empty_bar[slot], e_parity, "loader waiting for a free slot");
                } else if (lane == 0) {
                    int spins = 0;
                    while (released <= ls - p.S) {
                        bool progress = false;
                        const int c = scan_upto + 1, cs = c & (kCtr - 1);
                        if (c < n_load && ld_vol(&h->scan_seq[cs]) == c + 1) {
                            scan_upto = c;
                            progress = true;
                        }
                        const int d = tested_upto + 1, ds = d & (kCtr - 1);
                        if (d <= scan_upto && ld_vol(&h->tested[ds]) == ld_vol(&h->pushed[ds])) {
                            tested_upto = d;
                            progress = true;
                        }
                        // granule `released` is read by the scans of the next hg granules of its image (their rows above)
                        // and by the records of the granules before it (their row below)
                        // (tested_upto >= released - 1 also keeps the counter entries, reused every kCtr granules, apart)
                        if (scan_upto >= min(released + p.hg, rel_end - 1) && tested_upto >= released - 1) {
                            if (++released == rel_end) rel_end = min(n_load, rel_end + p.gpi);
                            progress = true;
                        }
                        if (!progress) {
                            __nanosleep(20);
                            SPIN_GUARD(spins, "loader waiting for a tested granule");
                        }
                    }
                }
                __syncwarp();
                STAT_ACC(13);
            }
            if (p.ring && lane == 0) {   // counters of the new granule (the entry's previous user, kCtr granules ago, is long done)
                h->pushed[ls & (kCtr - 1)] = 0;
                h->tested[ls & (kCtr - 1)] = 0;
                order_only();
            }
            const int npx = min(p.T, p.HW - gi * p.T);
            const float* src = p.yp + ((size_t)l_img * p.HW + (size_t)gi * p.T) * stride;
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
            if (++slot == p.S) {
                slot = 0;
                if (ls >= p.S) e_parity ^= 1u;
            }
            if (++gi == p.gpi) {
                gi = 0;
                ++l_img;
            }
        }
#ifdef CVM_DECODE_STATS
        STAT_ADD(14, clock64() - lo_t0);
        if (lane == 0) g_decode_cta[1][blockIdx.x] = gtime();
        stats_role_done(p);
#endif
        return;
    }
    if (warp < kTestWarps) tester_main<STRIDE, HM>(p, g0, lead, img0, n_segs);
    else if (warp - kTestWarps < p.S) scanner_main<STRIDE, HM, SEG>(p, warp - kTestWarps, g0, n_local, lead, n_load, img0, n_segs);
    else return;
#ifdef CVM_DECODE_STATS
    if (lane == 0) atomicMax(&g_decode_cta[warp < kTestWarps ? 3 : 2][blockIdx.x], gtime());
    stats_role_done(p);
#endif
}

// images whose predicted threshold did not hold and that the merge kernel recomputed the slow way (monitoring: a steady
// stream of these means the batches are too unlike each other for the prediction to pay)
__device__ unsigned long long g_decode_fallbacks;

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
    const SegMeta* segmeta;
    float* scores;
    int32_t* cls;
    long long* flat;
    float* centers;
    float* boxes;
    float* track;
    // first / last image of every scan CTA (host-computed, grid <= kMergeTab): the merge finds the CTAs of its image by
    // counting instead of 64-bit divisions (~2 us of serial latency at the top of each of these tiny CTAs)
    int use_tab;
    int first_img[kMergeTab], last_img[kMergeTab];
};

__global__ void __launch_bounds__(kMergeThreads) decode_merge_kernel(const __grid_constant__ MergeParams p) {
    __shared__ unsigned int hist[256];
    __shared__ int s_misc[4];
    __shared__ unsigned long long s_thr;
    __shared__ int s_cnt[kMergeThreads], s_off[kMergeThreads], s_g[2];
    __shared__ unsigned int s_need;
    __shared__ int s_total;

    const int tid = threadIdx.x, b = blockIdx.x, K = p.K;
    unsigned long long* const all = reinterpret_cast<unsigned long long*>(g_smem);   // [kMergeCap + K]
    unsigned long long* const keep = all + kMergeCap + K;                              // [K]
    unsigned long long* const sorted = keep + K;                                       // [K]

    // the scan CTAs whose granule range [g*n/G, (g+1)*n/G) overlaps this image's granules [lo, hi) (every CTA has at least
    // one granule: the grid never exceeds their number)
    const long long n_ch = p.n_steps, G = p.grid;
    long long g_first, g_last;
    if (p.use_tab) {   // the CTAs' image ranges are monotone: count the ones that end before / start at or before this image
        if (tid == 0) {
            s_need = 0u;
            s_total = 0;
        }
        g_first = __syncthreads_count(tid < p.grid && p.last_img[tid < p.grid ? tid : 0] < b);
        g_last = __syncthreads_count(tid < p.grid && p.first_img[tid < p.grid ? tid : 0] <= b) - 1;
    } else {
        if (tid == 0) {   // (64-bit divisions: once per CTA)
            const long long lo = (long long)b * p.spi, hi = lo + p.spi;
            long long gf = lo * G / n_ch;
            while (gf + 1 < G && (gf + 1) * n_ch / G <= lo) ++gf;
            long long gl = gf;
            while (gl + 1 < G && (gl + 1) * n_ch / G < hi) ++gl;
            s_g[0] = (int)gf;
            s_g[1] = (int)gl;
            s_need = 0u;
            s_total = 0;
        }
        group_sync(kMergeThreads);
        g_first = s_g[0];
        g_last = s_g[1];
    }
    const int n_over = (int)(g_last - g_first + 1);
    // Segments that ran behind a PROVISIONAL threshold (predicted from the image before, see flush_segment) published
    // every peak at or above it; the prediction holds for this image iff at least K of all its published peaks reach the
    // largest provisional threshold among its segments (then the K-th best score does, and nothing below a threshold can
    // belong to the top K).  One thread per segment fetches its record (one round trip for all of them).
    for (int t = tid; t < n_over; t += kMergeThreads) {
        const long long g = g_first + t;
        const int seg = b - (p.use_tab ? p.first_img[g] : (int)((g * n_ch / G) / p.spi));
        const SegMeta mt = p.segmeta[(size_t)g * p.max_segs + seg];
        if (t < kMergeThreads) {
            s_cnt[t] = mt.count;
            s_off[t] = (int)((size_t)g * p.max_segs + seg);
        }
        if (!mt.verified) atomicMax(&s_need, mt.thr_bits);
        atomicAdd(&s_total, mt.count);
    }
    group_sync(kMergeThreads);
    const unsigned need_bits = s_need;
    int n = 0, n_ok = 0;
    if (n_over <= 32 && s_total <= kMergeCap) {
        // the usual case: a few segments, everything fits: all keys in one sweep
        int base = 0;
        for (int t = 0; t < n_over; ++t) {
            const size_t o = (size_t)s_off[t];
            const int cnt = s_cnt[t];
            const unsigned long long* src = p.keys + o * p.seg_keys;
            for (int i = tid; i < cnt; i += kMergeThreads) {
                const unsigned long long k = src[i];
                all[base + i] = k;
                n_ok += (unsigned)(k >> 32) >= need_bits;
            }
            base += cnt;
        }
        n = base;
    } else {
        for (long long g = g_first; g <= g_last; ++g) {
            const int seg = b - (p.use_tab ? p.first_img[g] : (int)((g * n_ch / G) / p.spi));
            const size_t o = (size_t)g * p.max_segs + seg;
            const int cnt = p.segmeta[o].count;
            if (n + cnt > kMergeCap) {   // cnt <= seg_keys <= kMergeCap / 4, so n > K here
                group_sync(kMergeThreads);
                select_topk(all, n, K, hist, keep, s_misc, &s_thr, kMergeThreads);
                n = K;
            }
            const unsigned long long* src = p.keys + o * p.seg_keys;
            for (int i = tid; i < cnt; i += kMergeThreads) {
                const unsigned long long k = src[i];
                all[n + i] = k;
                n_ok += (unsigned)(k >> 32) >= need_bits;
            }
            n += cnt;
        }
    }
    if (need_bits != 0u) {
        if (tid == 0) s_misc[0] = 0;
        group_sync(kMergeThreads);
        n_ok = __reduce_add_sync(kFull, n_ok);
        if ((tid & 31) == 0 && n_ok) atomicAdd(&s_misc[0], n_ok);
        group_sync(kMergeThreads);
        n_ok = s_misc[0];
        group_sync(kMergeThreads);
        if (n_ok < K) {
            // The prediction was too high for this image: exact top-K of the whole image by this CTA alone (rare).  One
            // pixel per thread and step, channel by channel; a positive value above the running K-th key is tested against
            // its 3x3 neighbourhood in global memory; the buffer is cut back to its best K when it fills up.
            n = 0;
            if (tid == 0) {
                s_misc[3] = 0;
                s_thr = 0ull;
                atomicAdd(&g_decode_fallbacks, 1ull);
            }
            group_sync(kMergeThreads);
            const int HW = p.H * p.W;
            const float* img = p.yp + (size_t)b * HW * p.stride;
            const float ninf = __int_as_float(0xff800000);
            for (int pix0 = 0; pix0 < HW; pix0 += kMergeThreads) {
                const int pix = pix0 + tid;
                const int y = pix / p.W, x = pix - y * p.W;
                for (int c = 0; c < p.hm; ++c) {
                    if (pix < HW) {
                        const float* q = img + (size_t)pix * p.stride + c;
                        const float v = *q;
                        const unsigned long long key = ((unsigned long long)__float_as_uint(v) << 32) |
                                                       (unsigned long long)(0xFFFFFFFFu - ((unsigned)pix * (unsigned)p.hm + (unsigned)c));
                        if (v > 0.f && key > *(volatile unsigned long long*)&s_thr) {
                            float m = ninf;
                            for (int dy = -1; dy <= 1; ++dy)
                                for (int dx = -1; dx <= 1; ++dx) {
                                    if ((dy | dx) == 0 || y + dy < 0 || y + dy >= p.H || x + dx < 0 || x + dx >= p.W) continue;
                                    m = fmaxf(m, q[(dy * p.W + dx) * p.stride]);
                                }
                            if (m <= v) all[atomicAdd(&s_misc[3], 1)] = key;
                        }
                    }
                    group_sync(kMergeThreads);
                    const int cnt = *(volatile int*)&s_misc[3];
                    group_sync(kMergeThreads);
                    if (cnt > kMergeCap - kMergeThreads) {   // (room for one more step of appends is kept)
                        select_topk(all, cnt, K, hist, keep, s_misc, &s_thr, kMergeThreads);   // leaves s_thr = K-th key
                        if (tid == 0) s_misc[3] = K;
                        group_sync(kMergeThreads);
                    }
                }
            }
            n = *(volatile int*)&s_misc[3];
        }
    }
    group_sync(kMergeThreads);
    if (n > 1024) {   // many keys: radix select first
        select_topk(all, n, K, hist, keep, s_misc, &s_thr, kMergeThreads);
        n = K;
    }
    // rank by counting (keys are distinct): the K best in descending order; no barriers (the segments publish little more
    // than their best K keys, so n is a few hundred at most here)
    for (int i = tid; i < n; i += kMergeThreads) {
        const unsigned long long k = all[i];
        int rank = 0;
        for (int j = 0; j < n; ++j) rank += all[j] > k;
        if (rank < K) sorted[rank] = k;
    }
    if (n > K) n = K;
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



// SMs the scan kernel leaves free (cvm_decode_set_spare_sms): a data-parallel caller that issues its small collective just
// before the decode sets 1, so that the collective's kernel finds an SM at once instead of waiting for a persistent scan CTA
// to retire (every scan CTA takes a whole SM's shared memory)
int g_spare_sms = 0;

struct Plan {
    int T, gpi, S, ring, hg, gran_floats, cap, compact_at, grid, max_segs, seg_keys;
    long long n_gran;
    size_t smem_scan, smem_merge, ws_keys, ws_total;
};

int plan_decode(const cvm_layout* L, int stride, int B, int K, int spare_sms, Plan* t) {
    const int W = L->W;
    const long long HW = (long long)L->H * W;
    t->compact_at = K + kSlack;
    // a tester warp appends at most 32 keys between two looks at the compaction flag; four rounds of margin per warp
    t->cap = t->compact_at + 1 + 4 * kTestThreads;
    const size_t fixed = smem_bytes(0, 0, t->cap, K) + 32;
    if (fixed + 2 * 32 * (size_t)stride * 4 > (size_t)kSmemBudget) return CVM_ERR_ARG;
    const size_t px_bytes = (size_t)stride * 4;
    // Ring mode (the 3x3 neighbours of a record are read from shared memory): all kMaxSlots slots, each a whole number of
    // image rows (then a pixel's neighbourhood lies in the granules before / after its own: hg = 1), or, for rows wider than
    // a slot, 32-pixel multiples with hg = ceil((W + 1) / T) granules of history and lookahead (at most 3: the ring must
    // hold hg + 1 + hg granules plus the ones in flight).
    t->ring = 0;
    t->hg = 0;
    t->S = kMaxSlots;
    int T = 0;
    if (HW < (1 << 24)) {
        // whole rows: all slots, or - only in the kernel instantiation for 16-float pixels with 10 heatmap channels - one fewer
        for (int S = kMaxSlots; S >= (stride == 16 && L->hm == 10 ? kMaxSlots - 1 : kMaxSlots) && !t->ring; --S) {
            const long long px_slot = (long long)(((size_t)kSmemBudget - fixed) / S / px_bytes);
            if (px_slot < W) continue;
            long long rows = px_slot / W;
            if (rows > L->H) rows = L->H;
            while (rows > 1 && rows * W * (long long)px_bytes > kGranBytes) --rows;   // keep granules <= 32 KB ...
            while (rows > 1 && (L->H + rows - 1) / rows * (long long)B < 4LL * cvm_num_sms()) --rows;   // ... and a few per CTA
            while (rows > 1 && rows * W > kMaxT) --rows;
            if (rows * W > kMaxT) continue;
            T = (int)(rows * W);
            t->ring = 1;
            t->hg = 1;
            t->S = S;
        }
        const long long px_slot = (long long)(((size_t)kSmemBudget - fixed) / kMaxSlots / px_bytes);
        if (!t->ring && px_slot >= 32) {
            int T2 = (int)(px_slot / 32 * 32);
            if (T2 > kMaxT) T2 = kMaxT;
            const int hg = (W + 1 + T2 - 1) / T2;
            if (hg <= 3) {
                T = T2;
                t->ring = 1;
                t->hg = hg;
                t->S = kMaxSlots;
            }
        }
    }
    if (!t->ring) {
        // L2 mode: granules of <= 32 KB (a multiple of 32 pixels), one ring slot per scanner warp: all kScanWarps of them
        // when 32-pixel granules fit, fewer for very wide pixels
        T = (int)(kGranBytes / px_bytes / 32 * 32);
        if (T > kMaxT) T = kMaxT;
        if (T < 32) T = 32;
        if (HW < T) T = (int)((HW + 31) / 32 * 32);
        int S = kMaxSlots;
        while (S > 2 && fixed + (size_t)S * 32 * px_bytes > (size_t)kSmemBudget) --S;
        for (;; T -= 32) {
            if (T < 32) return CVM_ERR_ARG;
            if (fixed + (size_t)S * T * px_bytes <= (size_t)kSmemBudget) break;
        }
        t->S = S;
    }
    t->T = T;
    t->gran_floats = T * stride;
    t->gpi = (int)((HW + T - 1) / T);
    t->n_gran = (long long)B * t->gpi;
    // spare_sms: SMs left free, so that a small collective issued just before the decode by a data-parallel caller finds
    // an SM and runs beside it
    long long grid = cvm_num_sms() - spare_sms;
    if (grid > t->n_gran) grid = t->n_gran;
    if (grid < 1) grid = 1;
    t->grid = (int)grid;
    const long long len = (t->n_gran + grid - 1) / grid;
    t->max_segs = (int)((len - 1) / t->gpi + 2);
    t->smem_scan = smem_bytes(t->S, t->gran_floats, t->cap, K);
    t->smem_merge = ((size_t)kMergeCap + 3 * (size_t)K) * 8;
    t->seg_keys = K > kSegKeys ? K : kSegKeys;
    t->ws_keys = (size_t)grid * t->max_segs * t->seg_keys * 8;
    t->ws_total = t->ws_keys + (size_t)grid * t->max_segs * sizeof(SegMeta) + 16;
    return CVM_OK;
}

int check_decode_args(const cvm_layout* L, int stride, int B, int K) {
    CVM_CHECK_ARG(L != nullptr, "layout is NULL");
    CVM_CHECK_ARG(L->H > 0 && L->W > 0 && B >= 0, "bad shape");
    CVM_CHECK_ARG(L->hm >= 1 && L->hm <= 64 && stride >= L->Cp && L->Cp >= L->hm, "bad channel layout");
    CVM_CHECK_ARG(stride <= 256, "pixel stride above 256 floats is not supported");
    CVM_CHECK_ARG(K >= 1 && K <= kMaxK, "K=%d outside [1,%d]", K, kMaxK);
    CVM_CHECK_ARG((long long)L->H * L->W * L->hm < 0xFFFFFFFFLL, "H*W*hm must fit in 32 bits");
    CVM_CHECK_ARG((long long)L->H * L->W <= (long long)kRecPixel, "H*W above 2^29 - 1 is not supported");
    return CVM_OK;
}

template <int STRIDE, int HM, bool SEG, bool BULK>
int launch_scan_one(const DecodeParams& p, const Plan& t, cudaStream_t st) {
    CVM_SMEM_ATTR_ONCE((decode_scan_kernel<STRIDE, HM, SEG, BULK>), kSmemBudget);
    decode_scan_kernel<STRIDE, HM, SEG, BULK><<<t.grid, kThreads, t.smem_scan, st>>>(p);
    CVM_CHECK_LAUNCH("decode_scan_kernel");
    return CVM_OK;
}

template <int STRIDE, int HM>
int launch_scan(const DecodeParams& p, const Plan& t, bool bulk, cudaStream_t st) {
    if (p.seg_out)
        return bulk ? launch_scan_one<STRIDE, HM, true, true>(p, t, st) : launch_scan_one<STRIDE, HM, true, false>(p, t, st);
    return bulk ? launch_scan_one<STRIDE, HM, false, true>(p, t, st) : launch_scan_one<STRIDE, HM, false, false>(p, t, st);
}

int decode_impl(const cvm_layout* L, const float* y_pred, int pred_stride, int B, int K, const cvm_roi* rois, float* scores,
                int32_t* cls, long long* flat, float* centers, float* boxes, float* track, int seg_off, int seg_n,
                unsigned char* seg_out, void* ws, size_t ws_bytes, void* stream) {
    int rc = check_decode_args(L, pred_stride, B, K);
    if (rc != CVM_OK) return rc;
    CVM_CHECK_ARG(y_pred && scores && cls && flat && centers && boxes && ws, "NULL pointer argument");
    if (seg_out) CVM_CHECK_ARG(seg_n >= 1 && seg_n <= 256 && seg_off >= 0 && seg_off + seg_n <= pred_stride, "bad semseg slice");
    if (B == 0) return CVM_OK;
    Plan t;
    rc = plan_decode(L, pred_stride, B, K, g_spare_sms, &t);
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
    p.gpi = t.gpi;
    p.n_gran = t.n_gran;
    p.S = t.S;
    p.gran_floats = t.gran_floats;
    p.cap = t.cap;
    p.compact_at = t.compact_at;
    p.max_segs = t.max_segs;
    p.seg_keys = t.seg_keys;
    p.keys = static_cast<unsigned long long*>(ws);
    p.segmeta = reinterpret_cast<SegMeta*>(static_cast<unsigned char*>(ws) + t.ws_keys);
    p.hint = reinterpret_cast<unsigned long long*>(p.segmeta + (size_t)t.grid * t.max_segs);
    // the prediction is only taken from a call with the same map geometry and K
    p.cookie = 0x63766d6864656331ull ^ ((unsigned long long)L->H << 48) ^ ((unsigned long long)L->W << 32) ^
               ((unsigned long long)L->hm << 24) ^ ((unsigned long long)pred_stride << 12) ^ (unsigned long long)K;
    p.rescan_step = K / 2 > 8 ? K / 2 : 8;
#ifdef CVM_EXPERIMENT
    if (const char* e = getenv("CVM_DECODE_THR0")) {   // experiment only (results are wrong): start every segment at this score
        const float f = (float)atof(e);
        memcpy(&p.thr0_bits, &f, 4);
    }
#endif
    p.ring = t.ring;
    p.hg = t.hg;
    p.inv_T = 1.0f / (float)t.T;
    p.inv_W = 1.0f / (float)L->W;
    p.seg_out = seg_out;
    p.seg_off = seg_off;
    p.seg_n = seg_n;

    // the bulk-copy engine needs 16-byte granules: base pointer aligned and every image a whole number of them (full
    // granules are T*stride*4 bytes with T % 32 == 0, the partial last granule of an image then ends on one too)
    const bool bulk = cvm_aligned16(y_pred) && (((long long)p.HW * pred_stride) % 4 == 0) && (((long long)t.T * pred_stride) % 4 == 0);
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
    m.spi = t.gpi;
    m.grid = t.grid;
    m.max_segs = t.max_segs;
    m.seg_keys = t.seg_keys;
    m.n_steps = t.n_gran;
    m.off_roff = L->off_roff;
    m.off_box = L->off_box;
    m.off_track = track ? L->off_track : -1;
    m.off_class = L->off_class;
    m.nb_classes = L->nb_classes;
    m.R = (float)L->R;
    m.rois = rois;
    m.keys = p.keys;
    m.segmeta = p.segmeta;
    m.scores = scores;
    m.cls = cls;
    m.flat = flat;
    m.centers = centers;
    m.boxes = boxes;
    m.track = track;
    m.use_tab = t.grid <= kMergeTab;
    for (int g = 0; m.use_tab && g < t.grid; ++g) {   // the scan's split of the granules: CTA g owns [g*n/G, (g+1)*n/G)
        const long long lo = (long long)g * t.n_gran / t.grid, hi = (long long)(g + 1) * t.n_gran / t.grid - 1;
        m.first_img[g] = (int)(lo / t.gpi);
        m.last_img[g] = (int)(hi / t.gpi);
    }
    CVM_SMEM_ATTR_ONCE(decode_merge_kernel, ((size_t)kMergeCap + 3 * (size_t)kMaxK) * 8);
    // (tried: programmatic dependent launch of the merge kernel, griddepcontrol.launch_dependents at the top of the scan:
    // 0.143 ms instead of 0.139 for the pair - the early-resident merge CTAs are in the way more than the launch gap costs)
    decode_merge_kernel<<<B, kMergeThreads, t.smem_merge, st>>>(m);
    CVM_CHECK_LAUNCH("decode_merge_kernel");
    return CVM_OK;
}

}  // namespace

extern "C" size_t cvm_decode_topk_workspace_bytes(const cvm_layout* L, int pred_stride, int B, int K) {
    if (check_decode_args(L, pred_stride, B, K) != CVM_OK) return 0;
    Plan t;
    if (plan_decode(L, pred_stride, B > 0 ? B : 1, K, g_spare_sms, &t) != CVM_OK) return 0;
    return t.ws_total;
}

extern "C" int cvm_decode_topk(const cvm_layout* L, const float* y_pred, int pred_stride, int B, int K,
                               const cvm_roi* rois, float* scores, int32_t* cls, long long* flat, float* centers,
                               float* boxes, float* track, void* ws, size_t ws_bytes, void* stream) {
    return decode_impl(L, y_pred, pred_stride, B, K, rois, scores, cls, flat, centers, boxes, track, 0, 0, nullptr, ws, ws_bytes, stream);
}

// the tiling cvm_decode_topk would use: out[0..7] = pixels per granule, granules per image, ring slots, ring mode (1: 3x3
// neighbours from shared memory, 0: from global memory), halo granules, grid size, shared memory bytes, candidate capacity
extern "C" int cvm_decode_plan(const cvm_layout* L, int pred_stride, int B, int K, long long* out8) {
    int rc = check_decode_args(L, pred_stride, B, K);
    if (rc != CVM_OK) return rc;
    Plan t;
    rc = plan_decode(L, pred_stride, B > 0 ? B : 1, K, g_spare_sms, &t);
    if (rc != CVM_OK) return rc;
    out8[0] = t.T; out8[1] = t.gpi; out8[2] = t.S; out8[3] = t.ring; out8[4] = t.hg; out8[5] = t.grid;
    out8[6] = (long long)t.smem_scan; out8[7] = t.cap;
    return CVM_OK;
}

extern "C" int cvm_decode_set_spare_sms(int n) {
    CVM_CHECK_ARG(n >= 0 && n < 64, "spare SMs must be in [0, 64)");
    g_spare_sms = n;
    return CVM_OK;
}

extern "C" long long cvm_decode_fallback_count(void) {
    unsigned long long v = 0;
    if (cudaMemcpyFromSymbol(&v, g_decode_fallbacks, sizeof(v)) != cudaSuccess) return -1;
    return (long long)v;
}

extern "C" int cvm_decode_topk_semseg(const cvm_layout* L, const float* y_pred, int pred_stride, int B, int K,
                                      const cvm_roi* rois, float* scores, int32_t* cls, long long* flat, float* centers,
                                      float* boxes, float* track, int seg_off, int seg_n, unsigned char* seg_ids, void* ws,
                                      size_t ws_bytes, void* stream) {
    CVM_CHECK_ARG(seg_ids != nullptr, "seg_ids is NULL");
    return decode_impl(L, y_pred, pred_stride, B, K, rois, scores, cls, flat, centers, boxes, track, seg_off, seg_n, seg_ids, ws,
                       ws_bytes, stream);
}

#ifdef CVM_DECODE_STATS
extern "C" int cvm_decode_cta_times(unsigned long long* out /* [4][256] */) {
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(out, g_decode_cta, sizeof(unsigned long long) * 4 * 256);
    return 0;
}
extern "C" int cvm_decode_stats(unsigned long long* out32, int reset) {
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(out32, g_decode_stats, sizeof(unsigned long long) * 32);
    if (reset) {
        unsigned long long z[32] = {0};
        cudaMemcpyToSymbol(g_decode_stats, z, sizeof(z));
    }
    return 0;
}
#endif
