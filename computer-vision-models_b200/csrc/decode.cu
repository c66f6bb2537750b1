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
#include "common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kSlots = 4;         // row-segment ring depth
constexpr int kSlack = 256;       // buffer entries beyond K before a compaction (small: thresholds rise early)
constexpr int kMaxK = 1024;
constexpr int kE = 8;             // element-columns per thread
constexpr int kSlotFloats = 3328;  // ring slot size (13 KB), compile-time so that slot offsets are immediates

struct DecodeParams {
    const float* yp;
    long long total_floats;       // B*H*W*stride
    int stride, H, W, hm, K;
    int TW, SR, nbx, nsy;         // band width, stripe rows, bands per row, stripes per image
    int slot_floats;              // floats per ring slot (multiple of 4)
    int cap;                      // candidate buffer entries
    int compact_at;               // compact when more than this many candidates are buffered
    int use_bulk;
    float inv_hm;
    unsigned long long* keys;     // [B][NT][K]
    int* counts;                  // [B][NT]
};

// ---- exact top-K select on distinct 64-bit keys held in shared memory -------------------------------------------------
// On return keys[0..K) hold the K largest (unordered), *thr is the K-th largest key.  n > K required.  All threads call.
__device__ void select_topk(unsigned long long* keys, int n, int K, unsigned int* hist, unsigned long long* keep,
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
    unsigned long long kmin = ~0ull;
    for (int i = tid; i < n; i += kThreads) {
        const unsigned long long k = keys[i];
        if (k >= T) {
            keep[atomicAdd(&s_misc[3], 1)] = k;
            kmin = k < kmin ? k : kmin;
        }
    }
    __syncthreads();
    for (int i = tid; i < K; i += kThreads) keys[i] = keep[i];
    // exact threshold = smallest kept key (T may have zeroed low bits after an early exit)
    if (tid == 0) *thr_out = ~0ull;
    __syncthreads();
    atomicMin(thr_out, kmin);
    __syncthreads();
}

__device__ __forceinline__ float neg_inf() { return __int_as_float(0xff800000); }

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
    const float* gsrc;        // global address of (row ra-1, pixel px_lo, channel 0) (virtual when ra == 0)
    long long row_floats_g;   // W * stride
    int row_floats;           // (px_hi - px_lo) * stride
    int lead;                 // data of a row starts `lead` floats into its slot (16-byte granularity of the bulk copy)
    int r_last, ra, rb, H, K, compact_at, bulk;
    unsigned flat_x0, flat_row_step;
};

// global -> ring slot (row r lives in slot (r - (ra-1)) & 3): thread 0 with the bulk engine, or everybody with plain loads
__device__ __forceinline__ void load_row(const StreamCtx& c, int r, int tid) {
    const int i = r - (c.ra - 1), s = i & (kSlots - 1);
    float* dst = c.ring + (size_t)s * kSlotFloats;
    const float* src = c.gsrc + (long long)i * c.row_floats_g;       // first needed float
    if (c.bulk) {
        if (tid == 0) {
            const uint32_t bytes = (uint32_t)(((c.lead + c.row_floats + 3) & ~3) * 4);
            mbar_arrive_expect_tx(&c.full_bar[s], bytes);
            bulk_g2s(dst, src - c.lead, bytes, &c.full_bar[s]);
        }
    } else {
        for (int j = tid; j < c.row_floats; j += kThreads) dst[c.lead + j] = src[j];
    }
}

// One row: bring row r's values / 3-tap maxima into (vn, hn), then test row r-1 with (hp, hc, hn, vc).
// SLOT (= step index mod 4) and the slot size are compile-time constants: every shared load is [register + immediate].
template <int SLOT>
__device__ __forceinline__ void row_compute(const StreamCtx& c, int r, const float* const (&pv)[kE],
                                            const float* const (&pl)[kE], const float* const (&pr)[kE],
                                            const float (&hp)[kE], const float (&hc)[kE], float (&hn)[kE],
                                            const float (&vc)[kE], float (&vn)[kE], uint32_t& phase_bits, float thr_f,
                                            int& trigger) {
    if (r >= 0 && r < c.H) {
        if (c.bulk) {
            mbar_wait(&c.full_bar[SLOT], (phase_bits >> SLOT) & 1u);
            phase_bits ^= 1u << SLOT;
        } else {
            __syncthreads();
        }
        constexpr int so = SLOT * kSlotFloats;
#pragma unroll
        for (int k = 0; k < kE; ++k) {
            const float v = pv[k][so];
            vn[k] = v;
            hn[k] = fmaxf(v, fmaxf(pl[k][so], pr[k][so]));
        }
    } else {
#pragma unroll
        for (int k = 0; k < kE; ++k) vn[k] = hn[k] = neg_inf();
    }
    const int yt = r - 1;  // row whose 3x3 neighbourhood is now complete
    if (yt >= c.ra && yt < c.rb) {
        // hot path: ONE compare per thread and row against the running threshold score (positive by construction); only
        // threads holding a value that could still make the top K pay for the 3x3 tests.
        float vmax = vc[0];
#pragma unroll
        for (int k = 1; k < kE; ++k) vmax = fmaxf(vmax, vc[k]);
        if (vmax >= thr_f) {
            const unsigned flat_row = (unsigned)yt * c.flat_row_step + c.flat_x0;
            const unsigned cnt_addr = smem_u32(c.s_count), cand_addr = smem_u32(c.cand);
#pragma unroll
            for (int k = 0; k < kE; ++k) {
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
                }
            }
        }
    }
}

// End of a step of two rows: release the two slots, refill them, compact the candidate buffer when it passed the mark.
__device__ __forceinline__ void step_end(const StreamCtx& c, int r, int tid, float& thr_f, int& trigger) {
    // slots consumed by everyone; appends of these rows are visible.  The OR of the per-thread marks is the only race-free
    // uniform way to learn "buffer passed the mark" (fast threads may already append for the next rows once they leave
    // a plain barrier, so *s_count itself must not be sampled here).
    const int do_compact = __syncthreads_or(trigger);
    // refill with the rows 4 ahead (also after the virtual row -1 of a top stripe, whose slot is idle)
    if (r + kSlots >= 0 && r + kSlots <= c.r_last) load_row(c, r + kSlots, tid);
    if (r + 1 + kSlots <= c.r_last) load_row(c, r + 1 + kSlots, tid);
    if (do_compact) {  // every thread is in here, so *s_count is frozen
        select_topk(c.cand, *c.s_count, c.K, c.hist, c.keep, c.s_misc, c.s_thr);
        if (tid == 0) *c.s_count = c.K;
        trigger = 0;
        __syncthreads();
        thr_f = __uint_as_float((unsigned)(*c.s_thr >> 32));
    }
}

__global__ void __launch_bounds__(kThreads) decode_stream_kernel(const DecodeParams p) {
    static_assert(kSlots == 4, "the row-step rotation below has period 4");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ uint64_t full_bar[kSlots];
    __shared__ unsigned int hist[256];
    __shared__ int s_misc[4];
    __shared__ int s_count;
    __shared__ unsigned long long s_thr;

    StreamCtx c;
    c.ring = reinterpret_cast<float*>(smem_raw);
    c.cand = reinterpret_cast<unsigned long long*>(c.ring + (size_t)kSlots * kSlotFloats);
    c.keep = c.cand + p.cap;
    c.full_bar = full_bar;
    c.hist = hist;
    c.s_misc = s_misc;
    c.s_count = &s_count;
    c.s_thr = &s_thr;

    const int tid = threadIdx.x;
    int bid = blockIdx.x;
    const int bx = bid % p.nbx;
    bid /= p.nbx;
    const int sy = bid % p.nsy;
    const int b = bid / p.nsy;

    const int H = p.H, W = p.W, hm = p.hm, stride = p.stride;
    const int xa = bx * p.TW, xb = min(W, xa + p.TW);
    const int px_lo = max(xa - 1, 0), px_hi = min(xb + 1, W);
    c.ra = sy * p.SR;
    c.rb = min(H, c.ra + p.SR);
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
    // the bulk engine needs 16-byte granules: every row of this CTA must start at the same offset mod 4 floats, and the
    // rounded-up end of its last row must stay inside the tensor; otherwise this CTA uses plain cooperative loads
    {
        const long long f_hi_last = f_virtual + (long long)(c.r_last - (c.ra - 1)) * c.row_floats_g + c.row_floats;
        c.bulk = p.use_bulk && ((c.row_floats_g & 3LL) == 0) && (((f_hi_last + 3) & ~3LL) <= p.total_floats);
    }
    const int n_elem = (xb - xa) * hm;

    if (tid == 0) {
        for (int s = 0; s < kSlots; ++s) mbar_init(&full_bar[s], 1);
        mbar_fence_init();
        s_count = 0;
        s_thr = 0ull;
    }
    if (tid < kSlots * 4)  // -inf sentinels behind the data of every slot: neighbours outside the image, idle columns
        c.ring[(size_t)(tid >> 2) * kSlotFloats + kSlotFloats - 4 + (tid & 3)] = neg_inf();
    __syncthreads();

    for (int r = r_first; r <= c.r_last && r < c.ra - 1 + kSlots; ++r) load_row(c, r, tid);

    // fixed element-columns of this thread: addresses (inside slot 0) of the value and of its two x-neighbours
    const float* pv[kE];
    const float* pl[kE];
    const float* pr[kE];
    const float* const sent = c.ring + kSlotFloats - 4;
#pragma unroll
    for (int k = 0; k < kE; ++k) {
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
    float hA[kE], hB[kE], hC[kE], hD[kE], va[kE], vb[kE];
#pragma unroll
    for (int k = 0; k < kE; ++k) hA[k] = hB[k] = hC[k] = hD[k] = va[k] = vb[k] = neg_inf();

    uint32_t phase_bits = 0;
    float thr_f = __uint_as_float(1u);  // smallest positive float: "score > 0" and "score >= threshold" in one compare
    int trigger = 0;                    // this thread received a buffer position at/after the compaction mark

    // rows ra-1 .. rb, two per barrier.  Four h register sets rotate with period 4 (= ring depth, so the slot is a
    // compile-time constant), the two value sets with period 2: no register moves between rows.
#define CVM_ROW(J, HP, HC, HN, VC, VN) \
    row_compute<J>(c, r + J, pv, pl, pr, HP, HC, HN, VC, VN, phase_bits, thr_f, trigger)
    for (int r = c.ra - 1; r <= c.rb; r += 4) {
        CVM_ROW(0, hC, hD, hA, vb, va);
        if (r + 1 <= c.rb) CVM_ROW(1, hD, hA, hB, va, vb);
        step_end(c, r, tid, thr_f, trigger);
        if (r + 2 > c.rb) break;
        CVM_ROW(2, hA, hB, hC, vb, va);
        if (r + 3 <= c.rb) CVM_ROW(3, hB, hC, hD, va, vb);
        step_end(c, r + 2, tid, thr_f, trigger);
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
    int stride, H, W, hm, K, NT;
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
    __shared__ int s_n;
    __shared__ unsigned long long s_thr;
    __shared__ int s_off[257];

    unsigned long long* const all = reinterpret_cast<unsigned long long*>(smem_raw);   // [NT*K]
    unsigned long long* const keep = all + (size_t)p.NT * p.K;                          // [K]
    unsigned long long* const sorted = keep + p.K;                                      // [K]

    const int tid = threadIdx.x, b = blockIdx.x, K = p.K, NT = p.NT;
    // exclusive prefix of the tile counts (NT is small)
    if (tid == 0) {
        int acc = 0;
        for (int t = 0; t < NT; ++t) {
            if (t < 256) s_off[t] = acc;
            acc += p.counts[(size_t)b * NT + t];
        }
        s_n = acc;
    }
    __syncthreads();
    int n = s_n;
    for (int t = 0; t < NT; ++t) {
        const int cnt = p.counts[(size_t)b * NT + t];
        int off;
        if (t < 256) {
            off = s_off[t];
        } else {  // more than 256 tiles per image: recompute (never hit with the default tiling)
            off = 0;
            for (int u = 0; u < t; ++u) off += p.counts[(size_t)b * NT + u];
        }
        const unsigned long long* src = p.keys + ((size_t)b * NT + t) * K;
        for (int i = tid; i < cnt; i += kThreads) all[off + i] = src[i];
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
    int TW, SR, nbx, nsy, slot_floats, cap, compact_at;
    size_t smem_stream, smem_merge, ws_keys, ws_total;
};

int plan_tiling(const cvm_layout* L, int stride, int B, int K, Tiling* t) {
    const int hm = L->hm, H = L->H, W = L->W;
    int tw_max = (kThreads * kE) / hm;
    if (tw_max < 1) return CVM_ERR_ARG;
    // a row segment (band + 1-pixel halo each side + alignment slack + sentinels) must fit the fixed ring slot
    const int tw_slot = (kSlotFloats - 12) / stride - 2;
    if (tw_slot < 1) return CVM_ERR_ARG;
    if (tw_max > tw_slot) tw_max = tw_slot;
    t->nbx = (W + tw_max - 1) / tw_max;
    t->TW = (W + t->nbx - 1) / t->nbx;
    // stripes: enough CTAs to fill the machine ~3x, but never so many that the merge needs more than 16K keys
    const long long want = 3LL * cvm_num_sms() * 3;
    int nsy = (int)((want + (long long)B * t->nbx - 1) / ((long long)B * t->nbx));
    if (nsy < 1) nsy = 1;
    const int max_nsy_rows = (H + 7) / 8;  // at least 8 rows per stripe (halo overhead <= 25%)
    if (nsy > max_nsy_rows) nsy = max_nsy_rows;
    while (nsy > 1 && (long long)nsy * t->nbx * K > 16384) --nsy;
    if ((long long)nsy * t->nbx * K > 16384) return CVM_ERR_ARG;
    t->SR = (H + nsy - 1) / nsy;
    t->nsy = (H + t->SR - 1) / t->SR;
    t->slot_floats = kSlotFloats;
    t->compact_at = K + kSlack;
    t->cap = t->compact_at + 2 * t->TW * hm;   // two rows of appends between barriers
    t->smem_stream = (size_t)kSlots * t->slot_floats * 4 + (size_t)t->cap * 8 + (size_t)K * 8;
    const int NT = t->nbx * t->nsy;
    t->smem_merge = ((size_t)NT * K + 2 * (size_t)K) * 8;
    t->ws_keys = (size_t)B * NT * K * 8;
    t->ws_total = t->ws_keys + (size_t)B * NT * 4;
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
    const int NT = t.nbx * t.nsy;

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
    p.SR = t.SR;
    p.nbx = t.nbx;
    p.nsy = t.nsy;
    p.slot_floats = t.slot_floats;
    p.cap = t.cap;
    p.compact_at = t.compact_at;
    p.use_bulk = cvm_aligned16(y_pred);
    p.inv_hm = 1.0f / (float)L->hm;
    p.keys = static_cast<unsigned long long*>(ws);
    p.counts = reinterpret_cast<int*>(static_cast<unsigned char*>(ws) + t.ws_keys);

    CVM_CHECK_CUDA(cudaFuncSetAttribute(decode_stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)t.smem_stream));
    const long long grid = (long long)B * NT;
    CVM_CHECK_ARG(grid < 2147483647LL, "decode grid too large");
    decode_stream_kernel<<<(unsigned)grid, kThreads, t.smem_stream, st>>>(p);
    CVM_CHECK_LAUNCH("decode_stream_kernel");

    MergeParams m;
    memset(&m, 0, sizeof(m));
    m.yp = y_pred;
    m.stride = pred_stride;
    m.H = L->H;
    m.W = L->W;
    m.hm = L->hm;
    m.K = K;
    m.NT = NT;
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
