// Output decode, canonical CenterNet form (north_star): 3x3 peak NMS fused with an exact
// per-image top-K, then gather of the regression heads and box assembly (box math of the reference's
// models/centernet/post_processing.py:43-52 and common/utils/image.py:22-28, in fp32 like NumPy does it).
//
// Bound: HBM reads.  Algorithmic bytes per image: 4*H*W*pred_stride, read exactly once.
//
// Kernel 1 (decode_stream_kernel): one CTA per (image, row stripe, column band).  Thread 0 streams the band's row
// segments through a shared-memory ring with 1-D bulk async copies (UBLKCP + mbarrier); every thread owns E fixed
// (x, channel) element-columns and walks down the rows keeping the 3-tap horizontal maxima of the two previous rows in
// registers, so the 3x3 test costs 3 shared loads per element.  Survivors are turned into 64-bit keys
// (score bits << 32 | ~flat index): a total order with no ties, equal to (score desc, flat index asc).  Keys above the
// CTA's running threshold are appended to a shared-memory buffer; when it fills up an exact radix select (256-bin shared
// histogram per byte, warp-level suffix scan) keeps the best K and raises the threshold.  y_pred is never re-read.
// Kernel 2 (decode_merge_kernel): one CTA per image merges the <= K keys of each tile (same radix select), rank-sorts the
// K winners, fills a short tail with score-0 entries in flat-index order (tf.nn.top_k semantics on the masked map),
// gathers r_offset / fullbox / track_offset at the peaks and assembles boxes.
#include <stdlib.h>

#include "common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kSlots = 4;         // row-segment ring depth (2 rows consumed per step while 2 are in flight; deeper rings cost occupancy and measured slower)
constexpr int kSlack = 1024;      // buffer entries beyond K before the (rare) fallback compaction
constexpr int kScoreBins = 2048;  // score histogram: bin = float bits >> 20 (sign 0, 8 exponent bits, 3 mantissa bits)
constexpr int kMaxK = 1024;
// Two instantiations of the streaming kernel: <5, 2112> (narrow bands, ~80 registers, 3-4 CTAs/SM: the fast one) and
// <8, 3328> (wide pixel strides).  KE = element-columns per thread, SLOTF = floats per ring slot (compile-time so that
// slot offsets are immediates).

struct DecodeParams {
    const float* yp;
    long long total_floats;       // B*H*W*stride
    int stride, H, W, hm, K;
    int TW, nbx;                  // band width, bands per row
    int n_big, m_small, SR_small; // columns [0, n_big) are one tall tile each; the others are cut into m_small stripes
    int slot_floats;              // floats per ring slot (multiple of 4)
    int cap;                      // candidate buffer entries
    int compact_at;               // compact when more than this many candidates are buffered
    int use_bulk;
    float inv_hm;
    unsigned long long* keys;     // [n_tiles][K], tile = blockIdx.x
    int* counts;                  // [n_tiles]
};

__device__ __forceinline__ float neg_inf() { return __int_as_float(0xff800000); }

// ---- exact top-K select on distinct 64-bit keys held in shared memory -------------------------------------------------
// On return keys[0..K) hold the K largest (unordered), *thr is the K-th largest key.  n > K required.  All threads call.
__device__ __noinline__ void select_topk(unsigned long long* keys, int n, int K, unsigned int* hist, unsigned long long* keep,
                            int* s_misc /* [4] */, unsigned long long* thr_out) {
    const int tid = threadIdx.x;
    unsigned long long prefix = 0ull, mask = 0ull;
    int need = K;
    for (int shift = 56; shift >= 0; shift -= 8) {
        hist[tid] = 0;  // kThreads == 256 bins
        __syncthreads();
        for (int i = tid; i < n; i += kThreads) {
            const unsigned long long k = keys[i];
            if ((k & mask) == prefix) atomicAdd(&hist[(unsigned)(k >> shift) & 255u], 1u);
        }
        __syncthreads();
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
        __syncthreads();
        const int digit = s_misc[0];
        need = s_misc[1];
        const int pop = s_misc[2];
        prefix |= (unsigned long long)digit << shift;
        mask |= 0xFFull << shift;
        __syncthreads();  // s_misc is rewritten next pass
        if (pop == need) break;  // the whole bin is selected: every key >= prefix (low bits zero) is kept
    }
    const unsigned long long T = prefix;
    if (tid == 0) s_misc[3] = 0;
    __syncthreads();
    for (int i = tid; i < n; i += kThreads) {
        const unsigned long long k = keys[i];
        if (k >= T) keep[atomicAdd(&s_misc[3], 1)] = k;
    }
    __syncthreads();
    for (int i = tid; i < K; i += kThreads) keys[i] = keep[i];
    // T is the K-th largest key with (after an early exit) its undecided low bits zeroed: a valid, at most marginally
    // weaker, lower bound for every later candidate
    if (tid == 0) *thr_out = T;
    __syncthreads();
}

// Everything the per-row step needs that is not per-thread register state.
struct StreamCtx {
    float* ring;
    unsigned long long* cand;
    unsigned long long* keep;
    uint64_t* full_bar;
    unsigned int* hist;
    int* s_misc;
    int* s_count;
    unsigned long long* s_thr;
    unsigned int* shist;      // [kScoreBins] counts of buffered candidates per score bin (top 11 bits of the score)
    unsigned int* s_thr_bits; // running threshold score (float bits), raised by the histogram scan
    int* s_scanned;           // *s_count when the histogram was last scanned
    int* s_maxbin;            // highest score bin seen so far
    const float* gsrc;        // global address of (row ra-1, pixel px_lo, channel 0) (virtual when ra == 0)
    long long row_floats_g;   // W * stride
    int row_floats;           // (px_hi - px_lo) * stride
    int lead;                 // data of a row starts `lead` floats into its slot (16-byte granularity of the bulk copy)
    int r_last, ra, rb, H, K, compact_at, bulk;
    unsigned flat_x0, flat_row_step;
};

// global -> ring slot (row r lives in slot (r - (ra-1)) & 3): thread 0 with the bulk engine, or everybody with plain loads
template <int SLOTF, bool BULK>
__device__ __forceinline__ void load_row(const StreamCtx& c, int r, int tid) {
    const int i = r - (c.ra - 1), s = i & (kSlots - 1);
    float* dst = c.ring + (size_t)s * SLOTF;
    if (BULK) {
        if (tid == 0) {
            const float* src = c.gsrc + (long long)i * c.row_floats_g;   // first needed float
            const uint32_t bytes = (uint32_t)(((c.lead + c.row_floats + 3) & ~3) * 4);
            mbar_arrive_expect_tx(&c.full_bar[s], bytes);
            bulk_g2s(dst, src - c.lead, bytes, &c.full_bar[s]);
        }
    } else {
        const float* src = c.gsrc + (long long)i * c.row_floats_g;
        for (int j = tid; j < c.row_floats; j += kThreads) dst[c.lead + j] = src[j];
    }
}

// One row: bring row r's values / 3-tap maxima into (vn, hn), then test row r-1 with (hp, hc, hn, vc).
// SLOT (= step index mod 4) and the slot size are compile-time constants: every shared load is [register + immediate].
template <int SLOT, int KE, int SLOTF, bool BULK>
__device__ __forceinline__ void row_compute(const StreamCtx& c, int r, const float* const (&pv)[KE],
                                            const float* const (&pl)[KE], const float* const (&pr)[KE],
                                            const float (&hp)[KE], const float (&hc)[KE], float (&hn)[KE],
                                            const float (&vc)[KE], float (&vn)[KE], uint32_t& phase_bits, float thr_f,
                                            int& trigger) {
    if (BULK) {
        mbar_wait(&c.full_bar[SLOT], (phase_bits >> SLOT) & 1u);
        phase_bits ^= 1u << SLOT;
    } else {
        __syncthreads();
    }
    constexpr int so = SLOT * SLOTF;
#pragma unroll
    for (int k = 0; k < KE; ++k) {
        const float v = pv[k][so];
        vn[k] = v;
        hn[k] = fmaxf(v, fmaxf(pl[k][so], pr[k][so]));
    }
    const int yt = r - 1;  // row whose 3x3 neighbourhood is now complete
    if (yt >= c.ra && yt < c.rb) {
        // hot path: ONE compare per thread and row against the running threshold score (positive by construction); only
        // threads holding a value that could still make the top K pay for the 3x3 tests.
        float vmax = vc[0];
#pragma unroll
        for (int k = 1; k < KE; ++k) vmax = fmaxf(vmax, vc[k]);
        if (vmax >= thr_f) {
            const unsigned flat_row = (unsigned)yt * c.flat_row_step + c.flat_x0;
            const unsigned cnt_addr = smem_u32(c.s_count), cand_addr = smem_u32(c.cand), hist_addr = smem_u32(c.shist);
#pragma unroll
            for (int k = 0; k < KE; ++k) {
                // peak (value equals its 3x3 max) and not below the threshold score.  Scores equal to the threshold score
                // are appended without looking at the index half of the key: the next select drops them.
                if (vc[k] >= fmaxf(fmaxf(hp[k], hc[k]), fmaxf(hn[k], thr_f))) {
                    // a plain (not warp-aggregated) shared atomic: candidates are sparse, a short divergent path matters more
                    unsigned pos;
                    asm volatile("atom.shared.inc.u32 %0, [%1], 0x7fffffff;" : "=r"(pos) : "r"(cnt_addr) : "memory");
                    const unsigned flat = flat_row + (unsigned)(k * kThreads);
                    asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(cand_addr + pos * 8u), "r"(0xFFFFFFFFu - flat),
                                 "r"(__float_as_uint(vc[k]))
                                 : "memory");
                    trigger |= (pos >= (unsigned)c.compact_at);
                    // score histogram (exact for every bin at or above the running threshold): feeds the threshold scan
                    const unsigned bin = __float_as_uint(vc[k]) >> 20;
                    asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(hist_addr + bin * 4u) : "memory");
                    if ((int)bin > *(volatile int*)c.s_maxbin) atomicMax(c.s_maxbin, (int)bin);
                }
            }
        }
    }
}

// Threshold scan (one warp): the largest score bin tb with at least K buffered candidates in bins >= tb.  Bins at or above
// the running threshold are exact (everything that scores there was appended), so every later candidate below the
// lower edge of tb cannot make the top K: the edge becomes the new threshold.  No buffer traffic, no CTA-wide sync.
__device__ __noinline__ void scan_threshold(const unsigned int* shist, const int* s_count, int* s_scanned, const int* s_maxbin,
                                            unsigned int* s_thr_bits, int K, int lane) {
    const int n = *(volatile const int*)s_count;
    if (n == *s_scanned || n < K) return;
    const int top = *(volatile const int*)s_maxbin;
    unsigned above = 0;
    for (int base = top; base >= 0; base -= 32) {
        const int bin = base - lane;
        const unsigned h = bin >= 0 ? shist[bin] : 0u;
        unsigned incl = h;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        const unsigned hit = __ballot_sync(0xffffffffu, above + incl >= (unsigned)K);
        if (hit) {
            const int tb = base - (__ffs(hit) - 1);
            if (lane == 0) {
                const unsigned bits = (unsigned)tb << 20;
                if (bits > *s_thr_bits) *s_thr_bits = bits;
                *s_scanned = n;
            }
            return;
        }
        above += __shfl_sync(0xffffffffu, incl, 31);
    }
    if (lane == 0) *s_scanned = n;
}

// Fallback when the candidate buffer passed the mark: exact select, then the histogram is rebuilt from the K survivors.
__device__ __noinline__ void compact_buffer(unsigned long long* cand, int* s_count, int K, unsigned int* hist,
                                            unsigned long long* keep, int* s_misc, unsigned long long* s_thr,
                                            unsigned int* shist, int* s_scanned, unsigned int* s_thr_bits) {
    const int tid = threadIdx.x;
    select_topk(cand, *s_count, K, hist, keep, s_misc, s_thr);
    for (int i = tid; i < kScoreBins; i += kThreads) shist[i] = 0u;
    __syncthreads();
    for (int i = tid; i < K; i += kThreads) atomicAdd(&shist[(unsigned)(cand[i] >> 52)], 1u);
    if (tid == 0) {
        *s_count = K;
        *s_scanned = K;
        const unsigned bits = (unsigned)(*s_thr >> 32) & 0xFFF00000u;
        if (bits > *s_thr_bits) *s_thr_bits = bits;
    }
    __syncthreads();
}

// A bottom stripe ends with the virtual row H: its ring slot is filled with -inf and its barrier completed by hand, so
// the hot loop needs no special case.
template <int SLOTF, bool BULK>
__device__ __noinline__ void fill_virtual_row(float* ring, uint64_t* full_bar, int slot) {
    float* dst = ring + (size_t)slot * SLOTF;
    for (int j = threadIdx.x; j < SLOTF; j += kThreads) dst[j] = neg_inf();
    __syncthreads();
    if (BULK && threadIdx.x == 0) mbar_arrive_expect_tx(&full_bar[slot], 0);
}

// End of a step of two rows: release the two slots, refill them, raise the threshold from the score histogram; compact
// the candidate buffer only if it passed the mark (rare: the histogram threshold keeps the buffer short).
template <int SLOTF, bool BULK>
__device__ __forceinline__ void step_end(const StreamCtx& c, int r, int tid, float& thr_f, int& trigger) {
    // slots consumed by everyone; appends of these rows are visible.  The OR of the per-thread marks is the only race-free
    // uniform way to learn "buffer passed the mark" (fast threads may already append for the next rows once they leave
    // a plain barrier, so *s_count itself must not be sampled here).
    const int do_compact = __syncthreads_or(trigger);
    thr_f = __uint_as_float(*(volatile unsigned int*)c.s_thr_bits);   // written during the previous step: one step stale, still valid
    // refill with the rows kSlots ahead (also after the virtual row -1 of a top stripe, whose slot is idle)
#pragma unroll
    for (int d = 0; d < 2; ++d) {
        const int rr = r + d + kSlots;
        if (rr >= 0 && rr <= c.r_last)
            load_row<SLOTF, BULK>(c, rr, tid);
        else if (rr == c.H && c.rb == c.H)
            fill_virtual_row<SLOTF, BULK>(c.ring, c.full_bar, (rr - (c.ra - 1)) & (kSlots - 1));
    }
    if (do_compact) {  // every thread is in here, so *s_count is frozen
        compact_buffer(c.cand, c.s_count, c.K, c.hist, c.keep, c.s_misc, c.s_thr, c.shist, c.s_scanned, c.s_thr_bits);
        trigger = 0;
        thr_f = __uint_as_float(*(volatile unsigned int*)c.s_thr_bits);
    } else if ((tid >> 5) == 1) {
        scan_threshold(c.shist, c.s_count, c.s_scanned, c.s_maxbin, c.s_thr_bits, c.K, tid & 31);
    }
}

template <int KE, int SLOTF, int MIN_CTAS, bool BULK>
__global__ void __launch_bounds__(kThreads, MIN_CTAS) decode_stream_kernel(const DecodeParams p) {
    static_assert(kSlots == 4, "the row loop below is unrolled over the 4 ring slots");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ uint64_t full_bar[kSlots];
    __shared__ unsigned int hist[256];
    __shared__ int s_misc[4];
    __shared__ int s_count;
    __shared__ unsigned long long s_thr;
    __shared__ unsigned int s_thr_bits;
    __shared__ int s_scanned, s_maxbin;

    StreamCtx c;
    c.ring = reinterpret_cast<float*>(smem_raw);
    c.cand = reinterpret_cast<unsigned long long*>(c.ring + (size_t)kSlots * SLOTF);
    c.keep = c.cand + p.cap;
    c.shist = reinterpret_cast<unsigned int*>(c.keep + p.K);
    c.s_thr_bits = &s_thr_bits;
    c.s_scanned = &s_scanned;
    c.s_maxbin = &s_maxbin;
    c.full_bar = full_bar;
    c.hist = hist;
    c.s_misc = s_misc;
    c.s_count = &s_count;
    c.s_thr = &s_thr;

    const int tid = threadIdx.x;
    // tile geometry.  "Column" = (image, band).  The first n_big columns are processed top to bottom by one CTA each
    // (longest tiles first); the remaining ones are cut into m_small stripes so that the tail of the grid balances.
    int col, ra0, rb0;
    if ((int)blockIdx.x < p.n_big) {
        col = blockIdx.x;
        ra0 = 0;
        rb0 = p.H;
    } else {
        const int q = blockIdx.x - p.n_big;
        col = p.n_big + q / p.m_small;
        ra0 = (q % p.m_small) * p.SR_small;
        rb0 = min(p.H, ra0 + p.SR_small);
    }
    const int b = col / p.nbx, bx = col - b * p.nbx;

    const int H = p.H, W = p.W, hm = p.hm, stride = p.stride;
    const int xa = bx * p.TW, xb = min(W, xa + p.TW);
    const int px_lo = max(xa - 1, 0), px_hi = min(xb + 1, W);
    c.ra = ra0;
    c.rb = rb0;
    const int r_first = max(c.ra - 1, 0);
    c.r_last = min(c.rb, H - 1);
    c.H = H;
    c.K = p.K;
    c.compact_at = p.compact_at;
    c.row_floats = (px_hi - px_lo) * stride;
    c.row_floats_g = (long long)W * stride;
    const long long f_virtual = (((long long)b * H + (c.ra - 1)) * W + px_lo) * stride;   // may point before the tensor when ra == 0
    c.gsrc = p.yp + f_virtual;
    c.lead = (int)((f_virtual + (c.ra == 0 ? c.row_floats_g : 0)) & 3LL);
    c.flat_x0 = (unsigned)(xa * hm + tid);
    c.flat_row_step = (unsigned)(W * hm);
    c.bulk = BULK;
    const int n_elem = (xb - xa) * hm;

    if (tid == 0) {
        for (int s = 0; s < kSlots; ++s) mbar_init(&full_bar[s], 1);
        mbar_fence_init();
        s_count = 0;
        s_thr = 0ull;
        s_thr_bits = 1u;   // smallest positive float: "score > 0" and "score >= threshold" in one compare
        s_scanned = 0;
        s_maxbin = 0;
    }
    for (int i = tid; i < kScoreBins; i += kThreads) c.shist[i] = 0u;
    if (tid < kSlots * 4)  // -inf sentinels behind the data of every slot: neighbours outside the image, idle columns
        c.ring[(size_t)(tid >> 2) * SLOTF + SLOTF - 4 + (tid & 3)] = neg_inf();
    __syncthreads();

    for (int r = r_first; r < c.ra - 1 + kSlots; ++r) {
        if (r <= c.r_last)
            load_row<SLOTF, BULK>(c, r, tid);
        else if (r == H && c.rb == H)
            fill_virtual_row<SLOTF, BULK>(c.ring, full_bar, (r - (c.ra - 1)) & (kSlots - 1));
    }

    // fixed element-columns of this thread: addresses (inside slot 0) of the value and of its two x-neighbours
    const float* pv[KE];
    const float* pl[KE];
    const float* pr[KE];
    const float* const sent = c.ring + SLOTF - 4;
#pragma unroll
    for (int k = 0; k < KE; ++k) {
        const int e = tid + k * kThreads;
        pv[k] = pl[k] = pr[k] = sent;
        if (e < n_elem) {
            const int xl = hm == 1 ? e : __float2int_rz(((float)e + 0.5f) * p.inv_hm);
            const int ch = e - xl * hm;
            const int x = xa + xl;
            pv[k] = c.ring + c.lead + (x - px_lo) * stride + ch;
            if (x > 0) pl[k] = pv[k] - stride;
            if (x < W - 1) pr[k] = pv[k] + stride;
        }
    }
    float hA[KE], hB[KE], hC[KE], hD[KE], va[KE], vb[KE];
#pragma unroll
    for (int k = 0; k < KE; ++k) hA[k] = hB[k] = hC[k] = hD[k] = va[k] = vb[k] = neg_inf();

    uint32_t phase_bits = 0;
    float thr_f = __uint_as_float(1u);  // smallest positive float: "score > 0" and "score >= threshold" in one compare
    int trigger = 0;                    // this thread received a buffer position at/after the compaction mark

    // rows ra-1 .. rb, two per barrier.  Four h register sets rotate with period 4 (= ring depth, so the slot is a
    // compile-time constant), the two value sets with period 2: no register moves between rows.
    // Row -1 of a top stripe is virtual: its register sets already hold -inf, so it is skipped.  Row H below a bottom
    // stripe is a ring slot filled with -inf (fill_virtual_row): the loop treats it like any other row.
#define CVM_ROW(J, HP, HC, HN, VC, VN) \
    row_compute<J, KE, SLOTF, BULK>(c, r + J, pv, pl, pr, HP, HC, HN, VC, VN, phase_bits, thr_f, trigger)
    for (int r = c.ra - 1; r <= c.rb; r += 4) {
        if (r >= 0) CVM_ROW(0, hC, hD, hA, vb, va);
        if (r + 1 <= c.rb) CVM_ROW(1, hD, hA, hB, va, vb);
        step_end<SLOTF, BULK>(c, r, tid, thr_f, trigger);
        if (r + 2 > c.rb) break;
        CVM_ROW(2, hA, hB, hC, vb, va);
        if (r + 3 <= c.rb) CVM_ROW(3, hB, hC, hD, va, vb);
        step_end<SLOTF, BULK>(c, r + 2, tid, thr_f, trigger);
    }
#undef CVM_ROW

    int n = s_count;
    if (n > p.K) {
        select_topk(c.cand, n, p.K, hist, c.keep, s_misc, &s_thr);
        n = p.K;
    }
    unsigned long long* out = p.keys + (size_t)blockIdx.x * p.K;
    for (int i = tid; i < n; i += kThreads) out[i] = c.cand[i];
    if (tid == 0) p.counts[blockIdx.x] = n;
}

struct MergeParams {
    const float* yp;
    int stride, H, W, hm, K, nbx, n_big, m_small;
    int off_roff, off_box, off_track;
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

__global__ void __launch_bounds__(kThreads) decode_merge_kernel(const MergeParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ unsigned int hist[256];
    __shared__ int s_misc[4];
    __shared__ unsigned long long s_thr;

    const int tid = threadIdx.x, b = blockIdx.x, K = p.K;
    const int max_tiles = p.nbx * p.m_small;
    unsigned long long* const all = reinterpret_cast<unsigned long long*>(smem_raw);   // [max_tiles*K]
    unsigned long long* const keep = all + (size_t)max_tiles * K;                      // [K]
    unsigned long long* const sorted = keep + K;                                       // [K]

    // gather the keys of this image's tiles (a handful of tiles; walked by every thread, copied cooperatively)
    int n = 0;
    for (int bx = 0; bx < p.nbx; ++bx) {
        const int col = b * p.nbx + bx;
        const int t0 = col < p.n_big ? col : p.n_big + (col - p.n_big) * p.m_small;
        const int nt = col < p.n_big ? 1 : p.m_small;
        for (int t = t0; t < t0 + nt; ++t) {
            const int cnt = p.counts[t];
            const unsigned long long* src = p.keys + (size_t)t * K;
            for (int i = tid; i < cnt; i += kThreads) all[n + i] = src[i];
            n += cnt;
        }
    }
    __syncthreads();
    if (n > K) {
        select_topk(all, n, K, hist, keep, s_misc, &s_thr);
        n = K;
    }
    // rank sort (keys are distinct): descending
    for (int i = tid; i < n; i += kThreads) {
        const unsigned long long k = all[i];
        int rank = 0;
        for (int j = 0; j < n; ++j) rank += all[j] > k;
        sorted[rank] = k;
    }
    __syncthreads();

    const int hm = p.hm, W = p.W;
    const long long n_total = (long long)p.H * W * hm;
    const int K_eff = (long long)K < n_total ? K : (int)n_total;
    const cvm_roi roi = p.rois ? p.rois[b] : cvm_roi{1.0f, 0.0f, 0.0f, 0.0f};
    for (int i = tid; i < K; i += kThreads) {
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
                tx = __fadd_rn(cx, __fmul_rn(px[p.off_track], roi.inv_scale));
                ty = __fadd_rn(cy, __fmul_rn(px[p.off_track + 1], roi.inv_scale));
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

struct Tiling {
    int variant;   // 0: <5, 2112>, 1: <8, 3328>
    int KE, slot_floats;
    int TW, nbx, n_big, m_small, SR_small, n_tiles, cap, compact_at;
    size_t smem_stream, smem_merge, ws_keys, ws_total;
};

int plan_tiling(const cvm_layout* L, int stride, int B, int K, Tiling* t) {
    const int hm = L->hm, H = L->H, W = L->W;
    // variant 0 unless its ring slot cannot hold a useful band at this pixel stride
    t->variant = ((2112 - 12) / stride - 2 >= 16) ? 0 : 1;
    t->KE = t->variant == 0 ? 5 : 8;
    t->slot_floats = t->variant == 0 ? 2112 : 3328;
    int tw_max = (kThreads * t->KE) / hm;
    // a row segment (band + 1-pixel halo each side + alignment slack + sentinels) must fit the fixed ring slot
    const int tw_slot = (t->slot_floats - 12) / stride - 2;
    if (tw_max > tw_slot) tw_max = tw_slot;
    if (tw_max < 1) return CVM_ERR_ARG;
    t->nbx = (W + tw_max - 1) / tw_max;
    t->TW = (W + t->nbx - 1) / t->nbx;
    t->compact_at = K + kSlack;
    t->cap = t->compact_at + 2 * t->TW * hm;   // two rows of appends between barriers
    t->smem_stream = (size_t)kSlots * t->slot_floats * 4 + (size_t)t->cap * 8 + (size_t)K * 8 + (size_t)kScoreBins * 4;
    // Per-CTA fixed costs (threshold bootstrap, selects) favour tall tiles, whole waves favour many small ones: columns
    // (image x band) are processed top to bottom by one CTA each for as many whole waves as there are, the rest is cut
    // into m stripes (longest tiles first, short tiles fill the tail).
    int per_sm = (int)((220 * 1024) / (t->smem_stream + 2048));
    const int reg_cap = t->variant == 0 ? 3 : 1;   // CTAs/SM allowed by the register budget of each instantiation
    if (per_sm > reg_cap) per_sm = reg_cap;
    if (per_sm < 1) per_sm = 1;
    const long long slots = (long long)cvm_num_sms() * per_sm;
    const long long C = (long long)B * t->nbx;
    long long n_big = (C / slots) * slots;
    long long rem = C - n_big;
    int max_m = (H + 7) / 8;   // at least 8 rows per stripe (halo overhead <= 25%)
    if (max_m > 8) max_m = 8;
    while (max_m > 1 && (long long)max_m * t->nbx * K > 16384) --max_m;   // merge kernel: <= 16K keys per image
    int best_m = 1;
    double best_cost = 1e30;
    for (int m = 1; m <= max_m; ++m) {
        const long long waves = (rem * m + slots - 1) / slots;
        const double cost = (double)waves * (1.0 / m + 0.04);
        if (cost < best_cost - 1e-12) {
            best_cost = cost;
            best_m = m;
        }
    }
    if (const char* e = getenv("CVM_DECODE_NSY")) {  // experiment knob: uniform stripes
        const int v = atoi(e);
        if (v >= 1 && v <= (H + 7) / 8 && (long long)v * t->nbx * K <= 16384) {
            n_big = 0;
            rem = C;
            best_m = v;
        }
    }
    if ((long long)best_m * t->nbx * K > 16384) return CVM_ERR_ARG;
    t->n_big = (int)n_big;
    t->m_small = best_m;
    t->SR_small = (H + best_m - 1) / best_m;
    const long long n_tiles = n_big + rem * best_m;
    if (n_tiles >= 2147483647LL) return CVM_ERR_ARG;
    t->n_tiles = (int)n_tiles;
    t->smem_merge = ((size_t)t->nbx * best_m * K + 2 * (size_t)K) * 8;
    t->ws_keys = (size_t)n_tiles * K * 8;
    t->ws_total = t->ws_keys + (size_t)n_tiles * 4;
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

}  // namespace

extern "C" size_t cvm_decode_topk_workspace_bytes(const cvm_layout* L, int pred_stride, int B, int K) {
    if (check_decode_args(L, pred_stride, B, K) != CVM_OK) return 0;
    Tiling t;
    if (plan_tiling(L, pred_stride, B, K, &t) != CVM_OK) return 0;
    return t.ws_total;
}

extern "C" int cvm_decode_topk(const cvm_layout* L, const float* y_pred, int pred_stride, int B, int K,
                               const cvm_roi* rois, float* scores, int32_t* cls, long long* flat, float* centers,
                               float* boxes, float* track, void* ws, size_t ws_bytes, void* stream) {
    int rc = check_decode_args(L, pred_stride, B, K);
    if (rc != CVM_OK) return rc;
    CVM_CHECK_ARG(y_pred && scores && cls && flat && centers && boxes && ws, "NULL pointer argument");
    if (B == 0) return CVM_OK;
    Tiling t;
    rc = plan_tiling(L, pred_stride, B, K, &t);
    CVM_CHECK_ARG(rc == CVM_OK, "no tiling for H=%d W=%d hm=%d K=%d", L->H, L->W, L->hm, K);
    if (ws_bytes < t.ws_total) {
        cvm_set_error("workspace too small: %zu < %zu", ws_bytes, t.ws_total);
        return CVM_ERR_WS;
    }
    CVM_CHECK_ARG(t.smem_stream <= 200 * 1024 && t.smem_merge <= 200 * 1024, "shared memory budget exceeded");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    DecodeParams p;
    memset(&p, 0, sizeof(p));
    p.yp = y_pred;
    p.total_floats = (long long)B * L->H * L->W * pred_stride;
    p.stride = pred_stride;
    p.H = L->H;
    p.W = L->W;
    p.hm = L->hm;
    p.K = K;
    p.TW = t.TW;
    p.nbx = t.nbx;
    p.n_big = t.n_big;
    p.m_small = t.m_small;
    p.SR_small = t.SR_small;
    p.slot_floats = t.slot_floats;
    p.cap = t.cap;
    p.compact_at = t.compact_at;
    p.use_bulk = cvm_aligned16(y_pred);
    p.inv_hm = 1.0f / (float)L->hm;
    p.keys = static_cast<unsigned long long*>(ws);
    p.counts = reinterpret_cast<int*>(static_cast<unsigned char*>(ws) + t.ws_keys);

    const long long grid = t.n_tiles;
    // the bulk-copy engine needs 16-byte granules: base pointer aligned, every row starting at the same offset mod 4
    // floats (W*stride % 4 == 0) and the rounded-up end of the last row inside the tensor (total % 4 == 0)
    const bool bulk = cvm_aligned16(y_pred) && (((long long)L->W * pred_stride) % 4 == 0) && (p.total_floats % 4 == 0);
#define CVM_LAUNCH_STREAM(KE_, SLOTF_, MIN_, BULK_)                                                                           \
    do {                                                                                                                      \
        CVM_CHECK_CUDA(cudaFuncSetAttribute(decode_stream_kernel<KE_, SLOTF_, MIN_, BULK_>,                                   \
                                            cudaFuncAttributeMaxDynamicSharedMemorySize, (int)t.smem_stream));                \
        decode_stream_kernel<KE_, SLOTF_, MIN_, BULK_><<<(unsigned)grid, kThreads, t.smem_stream, st>>>(p);                   \
    } while (0)
    if (t.variant == 0) {
        if (bulk) CVM_LAUNCH_STREAM(5, 2112, 3, true);
        else CVM_LAUNCH_STREAM(5, 2112, 3, false);
    } else {
        if (bulk) CVM_LAUNCH_STREAM(8, 3328, 1, true);
        else CVM_LAUNCH_STREAM(8, 3328, 1, false);
    }
#undef CVM_LAUNCH_STREAM
    CVM_CHECK_LAUNCH("decode_stream_kernel");

    MergeParams m;
    memset(&m, 0, sizeof(m));
    m.yp = y_pred;
    m.stride = pred_stride;
    m.H = L->H;
    m.W = L->W;
    m.hm = L->hm;
    m.K = K;
    m.nbx = t.nbx;
    m.n_big = t.n_big;
    m.m_small = t.m_small;
    m.off_roff = L->off_roff;
    m.off_box = L->off_box;
    m.off_track = track ? L->off_track : -1;
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
    decode_merge_kernel<<<B, kThreads, t.smem_merge, st>>>(m);
    CVM_CHECK_LAUNCH("decode_merge_kernel");
    return CVM_OK;
}
