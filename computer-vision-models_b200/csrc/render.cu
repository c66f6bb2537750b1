// Ground-truth render: gaussian splat (max-combine) + loss-weights (min-combine) + centre scatter + ignore areas.
//
// Replaces fill_heatmap (reference models/centernet/processor.py:17-38, a numba double loop per object) and the render
// part of ProcessImages.process (processor.py:264-334) for a whole batch; also gen_prev_heatmap
// (models/centertracker/processor.py:22-41) with one plane and no weights plane.
//
// Bound: HBM writes.  Algorithmic bytes per image: 4*H*W*Ct, every byte written exactly once (zeros included).
//
// Structure: y_true is one flat list of chunks (P consecutive pixels of one image, all channels: a contiguous piece of
// the NHWC tensor, ~35 KB).  The list is cut into equal contiguous ranges, one persistent CTA per SM and range.  The CTA is
// warp-specialised: kGroups builder groups of 4 warps, each with its own staging buffer, take the chunks of the range
// round-robin and synchronise only inside the group (named barriers), and one setup group prepares the per-image state
// (derived object records, "last writer" flags of the centre scatter, ignore boxes, tables of the separable gaussian
// factors) one image ahead in a double-buffered table set.  A builder group builds a chunk in shared memory in its FINAL
// layout: pattern fill (zeros, ones in the weights channel); then every warp OWNS a contiguous quarter of the chunk's
// pixels and max-/min-combines into it, with plain loads and stores, every object whose window meets those pixels (found
// with a ballot over a compact window array, lane = object; then lane = column); then the regression targets at centre
// pixels and the ignore areas; then ONE bulk async copy (TMA engine: UBLKCP, shared -> global) streams the chunk out while
// the other groups build theirs, so the SM's store engine always has a chunk queued.  No store instruction touches global
// memory on the fast path.  All gaussian math is fp64 like the reference's (scalar fp64 stored as fp32).
//
// History (all measured on B200, 256 images of BASELINE configs[1], 755 MB): the first design (two 512-thread CTAs per SM,
// 80 KB chunks, CTA-wide barriers between the phases, one warp per (object, 32-column segment) unit combining with SHARED
// ATOMICS on the float bit patterns) ran at 0.162 ms.  Its limiter was not the barriers but the atomics: ATOMS retires
// 2 cycles per LANE (B300_MICROARCH.md), 20 M of them per launch = 0.14 ms of every SM's load/store unit.  Giving each
// warp exclusive pixels removes them: 0.218 ms (this structure with atomics) -> 0.148 -> 0.142 ms with the table / buffer
// sizes below.  Also measured: 3 / 4 / 6 groups (0.152-0.166), 96 / 160 threads per group (slower: pieces too large /
// register cap), cold paths out of line (__noinline__: 0.203 ms - the generic-address loads of the parameter block and
// the call ABI cost more than the instruction-cache footprint), a 128-byte object record (a 32-way bank conflict for
// lane = object reads: hence the 136-byte stride and the compact window array).  What is left is latency: ~100
// instructions per (warp, object) visit at ~8 cycles each with 20 builder warps per SM; a bare bulk-store stream of the
// same bytes takes 0.123 ms (tools/write_bench.cu).
#include <stdlib.h>

#include "common.cuh"

namespace {

// -DCVM_EXPERIMENT: CVM_RENDER_SKIP switches phases of the kernel off (tools/time_render.py ablations; results are wrong).
// The shipped build has no such knob: RDBG() is a compile-time 0.
#ifdef CVM_EXPERIMENT
#define RDBG(bit) (p.dbg_skip & (bit))
#else
#define RDBG(bit) 0
#endif
// Builder groups (see above): kGroups groups of kGT threads with a kBufBytes staging buffer each, plus one setup group
// of kST threads.  80 registers per thread at 704 threads (the allocation granularity leaves no room for more warps).
#ifndef CVM_RENDER_GROUPS
#define CVM_RENDER_GROUPS 5
#endif
#ifndef CVM_RENDER_GT
#define CVM_RENDER_GT 128
#endif
#ifndef CVM_RENDER_BUF
#define CVM_RENDER_BUF 34816
#endif
#ifndef CVM_RENDER_FACTAB
#define CVM_RENDER_FACTAB 2048   /* the 32 objects of BASELINE configs[1] need ~1200 entries (sd 150), the ~40 of configs[3] ~1500 */
#endif
constexpr int kGroups = CVM_RENDER_GROUPS;
constexpr int kGT = CVM_RENDER_GT;            // threads per group (>= kMaxObjSmem: one thread per object in the scatter)
constexpr int kGW = kGT / 32;                 // warps per group
constexpr int kST = 128;                       // threads of the setup group
constexpr int kThreads = kGroups * kGT + kST;
constexpr int kBufBytes = CVM_RENDER_BUF;     // staging buffer per builder group
constexpr int kMaxObjSmem = 64;    // objects with tabulated factors per image (later ones evaluate exp per cell)
constexpr int kMaxIgnSmem = 16;    // ignore boxes cached in shared memory per image (more are read from global)
constexpr int kFacTab = CVM_RENDER_FACTAB;    // entries of the per-image table of gaussian factors (objects that do not fit use exp)
static_assert(kGT >= kMaxObjSmem && kGT % 32 == 0 && kST >= kMaxObjSmem && kST % 32 == 0, "one thread per tabulated object");

struct RenderParams {
    const cvm_obj* objs;
    const int32_t* obj_offsets;
    const cvm_box* ignore;
    const int32_t* ign_offsets;
    float* out;        // [B,H,W,Cout]
    int H, W, Cout;
    int HW;
    int n_planes;      // heatmap planes (hm)
    int wch;           // weights channel in the output pixel, -1 = none (prev-frame heatmap)
    int P;             // pixels per chunk (multiple of 4)
    int cpi;           // chunks per image
    long long n_chunks;
    int per_class;     // 1: heat plane = obj.cls (Profile N); 0: plane 0 (reference as shipped)
    int force_explicit;
    int off_class, off_roff, off_box, off_track;
    const float* extra;  // per-object extra regression targets [n_obj][extra_stride] (l_shape / 3d_info, processor.py:296-299) or NULL
    int extra_stride, extra_off, extra_n;   // written to channels [extra_off, extra_off + extra_n) of the centre pixel
    int bulk;          // chunks are 16-byte aligned in global memory: stream them out with bulk async copies
    int vec;           // chunks are 16-byte aligned: the plain-store path may use 128-bit stores
    int dbg_skip;      // experiment knob (CVM_RENDER_SKIP): 1 = no splat, 2 = no store, 4 = no fill
    double R, alpha;
};

// derived per-object quantities, computed once per image by one thread per object
struct ObjDerived {
    double inv2vx, inv2vy;  // 1/(2*var)
    double rw;              // reduce_weight (processor.py:24)
    double peak;
    int cx, cy;             // gaussian centre (mask px; may be outside the map for explicit centres)
    int x0, x1, y0, y1;     // clamped half-open window (processor.py:26-29)
    int plane;
    int scx, scy;           // scatter pixel (clamped centre), -1 if no scatter
    float offx, offy, bw, bh, tx, ty;
    int cls;
    int last_at_pixel;      // no later object scatters to the same pixel
    int tab;                // start of the object's column factors in the table, -1 if they did not fit
    int tabr;               // start of the object's row factors in the table, -1 if they did not fit
    int lox, nx;            // the factors are even in the distance to the centre: entries for |x - cx| in [lox, lox + nx)
    int loy, ny;
    int pad[2];             // 136 bytes: a stride of 128 would put the same field of all objects into one shared-memory bank
};

__device__ __forceinline__ void derive(const cvm_obj& o, const RenderParams& p, ObjDerived& d) {
    const double w = o.w, h = o.h;
    int cx, cy;
    d.scx = -1;
    d.scy = -1;
    if (p.force_explicit || (o.flags & CVM_OBJ_EXPLICIT_CENTER)) {
        cx = o.cx;
        cy = o.cy;
        if (!(o.flags & CVM_OBJ_NO_SCATTER) && cx >= 0 && cx < p.W && cy >= 0 && cy < p.H) {
            d.scx = cx;
            d.scy = cy;
        }
        d.offx = d.offy = 0.f;
    } else {
        // calc_img_data, processor.py:60-67 (fp64, int() truncates toward zero)
        const double sx = o.x / p.R, sy = o.y / p.R, sw = o.w / p.R, sh = o.h / p.R;
        const double cxf = sx + sw / 2.0, cyf = sy + sh / 2.0;
        cx = max(0, min(p.W - 1, (int)cxf));
        cy = max(0, min(p.H - 1, (int)cyf));
        d.offx = (float)(cxf - cx);
        d.offy = (float)(cyf - cy);
        if (!(o.flags & CVM_OBJ_NO_SCATTER)) {
            d.scx = cx;
            d.scy = cy;
        }
    }
    d.cx = cx;
    d.cy = cy;
    d.bw = (float)w;
    d.bh = (float)h;
    d.tx = o.track[0];
    d.ty = o.track[1];
    d.cls = o.cls;
    d.peak = (double)o.peak;
    d.rw = 1.0 - ((fmin(20.0, fmax(w, h)) / 16.0) - 0.25);            // processor.py:23-24
    const int hx = (int)floor(w / 2.0), hy = (int)floor(h / 2.0);     // int(width // 2)
    d.x0 = max(0, cx - hx);
    d.x1 = min(p.W, cx + hx);
    d.y0 = max(0, cy - hy);
    d.y1 = min(p.H, cy + hy);
    const double sdx = (p.alpha * w) / (6.0 * p.R), sdy = (p.alpha * h) / (6.0 * p.R);
    d.inv2vx = 1.0 / (2.0 * (sdx * sdx));                             // processor.py:32-35
    d.inv2vy = 1.0 / (2.0 * (sdy * sdy));
    d.plane = p.per_class ? o.cls : 0;
    if (d.plane < 0 || d.plane >= p.n_planes) d.x1 = d.x0;            // class out of range: draw nothing
    d.last_at_pixel = 1;
    // distances |x - cx| that occur in the window (the centre of an explicit object may lie outside the map)
    d.lox = d.loy = d.nx = d.ny = 0;
    if (d.x0 < d.x1 && d.y0 < d.y1) {
        d.lox = cx < d.x0 ? d.x0 - cx : (cx >= d.x1 ? cx - (d.x1 - 1) : 0);
        d.nx = max(cx - d.x0, d.x1 - 1 - cx) - d.lox + 1;
        d.loy = cy < d.y0 ? d.y0 - cy : (cy >= d.y1 ? cy - (d.y1 - 1) : 0);
        d.ny = max(cy - d.y0, d.y1 - 1 - cy) - d.loy + 1;
    }
}

// shared -> global bulk copy (bytes % 16 == 0, both addresses 16-byte aligned), tracked by the issuing thread's bulk groups
__device__ __forceinline__ void bulk_s2g(void* dst_gmem, const void* src_smem, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem), "r"(smem_u32(src_smem)), "r"(bytes)
                 : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}

extern __shared__ __align__(128) unsigned char g_render_smem[];   // RenderShared, then the staging buffers of the groups
#ifdef CVM_EXPERIMENT
__device__ unsigned long long g_render_dbg[12];   // cycles of group leaders per phase (tools/time_render.py)
#define RCLK(i)                                  \
    {                                            \
        const long long now_ = clock64();        \
        dbg_t[i] += now_ - dbg_last;             \
        dbg_last = now_;                         \
    }
#else
#define RCLK(i)
#endif

// Per-image state, built by the setup group one image ahead of the builders (two sets, image parity picks one).
// The gaussian is separable: exp(-(ax + ay)) = exp(-ax) * exp(-ay).  The column factors exp(-ax) of every object
// and its row factors exp(-ay) are tabulated once per image, so a covered pixel costs one fp64 multiply instead
// of one fp64 exp.  The product differs from the exp of the sum by a few ulp of a DOUBLE; after the rounding to fp32
// that the reference stores, the value is the same except when a rounding boundary falls inside that interval
// (probability ~1e-8 per value), where it differs by one fp32 ulp -- far inside the 1e-5 tolerance.
struct TableSet {
    ObjDerived obj[kMaxObjSmem];
    int4 win[kMaxObjSmem];             // x0, x1, y0, y1 of every object again: what the builders' window test reads (lane = object)
    double fac[kFacTab];               // object after object: its column factors, then its row factors
    int ign[kMaxIgnSmem][4];           // ignore boxes of the image: sx, ex, sy, ey (clamped to the map)
    int o_begin, o_end, n, i_begin, n_ign, pad[3];
};

struct RenderShared {
    TableSet set[2];
    unsigned char own[kFacTab];             // (setup scratch) object that owns each table entry
    int used;                               // (setup scratch) table entries in use
    volatile int ready;                     // images of this CTA, counted from its first one, whose table set is complete
    volatile int prog[kGroups];             // image (same counting) each builder group is working in
};

__device__ __forceinline__ void bar_group(int id, int nt) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nt) : "memory"); }
__device__ __forceinline__ void fence_cta_r() { asm volatile("fence.acq_rel.cta;" ::: "memory"); }
#define RENDER_SPIN(cond)                          \
    {                                              \
        unsigned spins_ = 0;                       \
        while (!(cond)) {                          \
            __nanosleep(32);                       \
            if (++spins_ > (1u << 24)) __trap();   \
        }                                          \
    }

// ---- per-image tables: derived records, scatter winners, ignore boxes, gaussian factors.  Run by `nt` threads (st = 0 ..
//      nt - 1, nt >= kMaxObjSmem) that share the named barrier `bar_id` ----
__device__ __forceinline__ void setup_image(const RenderParams& p, RenderShared& S, TableSet& T, int img, int st, int nt, int bar_id) {
#define BAR() bar_group(bar_id, nt)
    const int W = p.W;
    const bool scatter = p.off_class >= 0 || p.off_roff >= 0 || p.off_box >= 0 || p.off_track >= 0 || p.extra_n > 0;
    const int o_begin = p.obj_offsets[img], o_end = p.obj_offsets[img + 1];
    const int n = min(kMaxObjSmem, o_end - o_begin);
    int n_ign = 0, i_begin = 0;
    if (p.ignore != nullptr && p.wch >= 0) {
        i_begin = p.ign_offsets[img];
        n_ign = p.ign_offsets[img + 1] - i_begin;
    }
    if (st < min(n_ign, kMaxIgnSmem)) {
        const cvm_box bx = p.ignore[i_begin + st];
        T.ign[st][0] = max((int)bx.x, 0);
        T.ign[st][1] = min(max((int)(bx.x + bx.w), 0), W);
        T.ign[st][2] = max((int)bx.y, 0);
        T.ign[st][3] = min(max((int)(bx.y + bx.h), 0), p.H);
    }
    if (st < n) {
        derive(p.objs[o_begin + st], p, T.obj[st]);
        T.win[st] = make_int4(T.obj[st].x0, T.obj[st].x1, T.obj[st].y0, T.obj[st].y1);
    }
    BAR();
    if (st < n && scatter) {
        // a later object (list order) on the same centre pixel overwrites r_offset / fullbox / track_offset
        // (processor.py:288-299), so only the last one writes them
        const int sx = T.obj[st].scx, sy = T.obj[st].scy;
        bool last = true;
        for (int j = st + 1; o_begin + j < o_end && last; ++j) {
            if (j < n) {
                if (T.obj[j].scx == sx && T.obj[j].scy == sy) last = false;
            } else {
                ObjDerived e;
                derive(p.objs[o_begin + j], p, e);
                if (e.scx == sx && e.scy == sy) last = false;
            }
        }
        T.obj[st].last_at_pixel = last;
    }
    if (st < n) {   // table space in object order; an object that does not fit evaluates exp per cell
        int off = 0;
        for (int o = 0; o < st; ++o) off += T.obj[o].nx + T.obj[o].ny;
        const int wd = T.obj[st].nx, ht = T.obj[st].ny;
        const bool fits = off + wd + ht <= kFacTab;   // (then every object before it fits too)
        // owner of every table entry, so that the factors can be computed one entry per thread
        if (fits)
            for (int e = 0; e < wd + ht; ++e) S.own[off + e] = (unsigned char)st;
        T.obj[st].tab = fits ? off : -1;
        T.obj[st].tabr = fits ? off + wd : -1;
        if (fits && (st == n - 1 || off + wd + ht + T.obj[st + 1].nx + T.obj[st + 1].ny > kFacTab)) S.used = off + wd + ht;
    }
    if (st == 0 && (n == 0 || T.obj[0].nx + T.obj[0].ny > kFacTab)) S.used = 0;
    BAR();
    const int used = S.used;
    for (int e = st; e < used; e += nt) {
        const ObjDerived& d = T.obj[S.own[e]];
        if (e < d.tabr) {
            const double dx = (double)(d.lox + (e - d.tab));
            T.fac[e] = exp(-(dx * dx * d.inv2vx));
        } else {
            const double dy = (double)(d.loy + (e - d.tabr));
            T.fac[e] = exp(-(dy * dy * d.inv2vy));
        }
    }
    if (st == 0) {
        T.o_begin = o_begin;
        T.o_end = o_end;
        T.n = n;
        T.i_begin = i_begin;
        T.n_ign = n_ign;
    }
#undef BAR
}

// ---- the setup group: the tables of the CTA's images, one image ahead of the builders (the first image was prepared by
//      the whole CTA) ----
__device__ __forceinline__ void setup_main(const RenderParams& p, RenderShared& S, int st, int img0, int n_img) {
    for (int k = 1; k < n_img; ++k) {
        if (k >= 2 && st == 0) {   // the set still holds image k - 2: every builder group must have left it
            for (int g = 0; g < kGroups; ++g) RENDER_SPIN(S.prog[g] >= k - 1);
            fence_cta_r();
        }
        bar_group(kGroups + 1, kST);
        setup_image(p, S, S.set[k & 1], img0 + k, st, kST, kGroups + 1);
        fence_cta_r();
        bar_group(kGroups + 1, kST);   // (also: S.own / S.used are free for the next image)
        if (st == 0) {
            fence_cta_r();   // release: the group's writes, observed through the barrier, before the flag
            S.ready = k + 1;
        }
    }
}

// One object into the pixels [s0, s1) of the image that ONE warp owns (a piece of the chunk that starts at pixel q0; rows
// ysa..ysb): lane = column.  Nobody else touches these cells during the splat, so the max / min combine is a plain
// read-modify-write - no shared atomics (they retire at 2 cycles per LANE on this SM and were the kernel's bottleneck).
// TAB: both factors come from the image's tables (col / row); otherwise exp per cell where a table is missing.
template <bool TAB>
__device__ __forceinline__ void splat_rows(const RenderParams& p, const ObjDerived& d, const double* col, const double* row,
                                           float* st, int q0, int s0, int s1, int ysa, int ysb, int lane) {
    const int x0 = d.x0, x1 = d.x1, cx = d.cx, cy = d.cy;
    const bool has_tab = TAB || (col && d.tab >= 0), has_tabr = TAB || (row && d.tabr >= 0);
    const int tab = d.tab - d.lox, tabr = d.tabr - d.loy;
    const int r0 = max(d.y0, ysa), r1 = min(d.y1, ysb + 1);
    const double peak = d.peak, rw = d.rw, inv2vx = d.inv2vx, inv2vy = d.inv2vy;
    const int W = p.W, Cout = p.Cout, wch = p.wch;
    float* const plane_px = st + d.plane;
    for (int y = r0; y < r1; ++y) {
        const int rs = y * W;
        const int cs = max(x0, s0 - rs), ce = min(x1, s1 - rs);   // the part of the window's row inside [s0, s1)
        if (cs >= ce) continue;
        double ey;
        if (has_tabr) {
            ey = row[tabr + abs(y - cy)];
        } else {
            const double dy = (double)(y - cy);
            ey = exp(-(dy * dy * inv2vy));
        }
        for (int x = cs + lane; x < ce; x += 32) {
            double ex;
            if (has_tab) {
                ex = col[tab + abs(x - cx)];
            } else {
                const double dx = (double)(x - cx);
                ex = exp(-(dx * dx * inv2vx));
            }
            const double gv = ex * ey;                                           // processor.py:34-36
            float* const cell = plane_px + (rs + x - q0) * Cout;
            *cell = fmaxf(*cell, (float)(gv * peak));                            // :37
            if (wch >= 0) {
                float* const wc = st + (rs + x - q0) * Cout + wch;
                *wc = fminf(*wc, (float)(1.0 - rw * gv));                        // :38
            }
        }
    }
    __syncwarp();   // the next object of this warp may meet the same cells from other lanes
}
__device__ __forceinline__ void splat_object_slow(const RenderParams& p, const ObjDerived& d, const double* col, const double* row,
                                               float* st, int q0, int s0, int s1, int ysa, int ysb, int lane) {
    splat_rows<false>(p, d, col, row, st, q0, s0, s1, ysa, ysb, lane);
}
__device__ __forceinline__ void splat_object(const RenderParams& p, const ObjDerived& d, const double* col, const double* row,
                                             float* st, int q0, int s0, int s1, int ysa, int ysb, int lane) {
    if (d.tab >= 0 && d.tabr >= 0) splat_rows<true>(p, d, col, row, st, q0, s0, s1, ysa, ysb, lane);
    else splat_object_slow(p, d, col, row, st, q0, s0, s1, ysa, ysb, lane);
}

// centre scatter of one object: regression targets and the class one-hot live in channels the splat never touches
__device__ __forceinline__ void scatter_object(const RenderParams& p, const ObjDerived& d, int obj_index, float* st, int q0,
                                               int q1) {
    const int qs = d.scy * p.W + d.scx;
    if (d.scx < 0 || qs < q0 || qs >= q1) return;
    float* px = st + (size_t)(qs - q0) * p.Cout;
    if (p.off_class >= 0 && d.cls >= 0 && p.off_class + d.cls < p.Cout) px[p.off_class + d.cls] = 1.0f;   // :290
    if (!d.last_at_pixel) return;
    if (p.off_roff >= 0) {
        px[p.off_roff] = d.offx;                                                                         // :292
        px[p.off_roff + 1] = d.offy;
    }
    if (p.off_box >= 0) {
        px[p.off_box] = d.bw;                                                                            // :294
        px[p.off_box + 1] = d.bh;
    }
    if (p.off_track >= 0) {
        px[p.off_track] = d.tx;
        px[p.off_track + 1] = d.ty;
    }
    if (p.extra_n > 0) {   // l_shape (7) / 3d_info (5) targets, computed on the host (processor.py:69-115,296-299)
        const float* e = p.extra + (size_t)obj_index * p.extra_stride;
        for (int k = 0; k < p.extra_n; ++k) px[p.extra_off + k] = e[k];
    }
}

// objects beyond the tabulated ones (crowded images; rare): derived on the fly, exp per cell; every warp walks all of them
// for its own pixels
__device__ __forceinline__ void splat_extra_objects(const RenderParams& p, float* st, int o_begin, int o_end, int q0, int q1, int s0,
                                                 int s1, int ysa, int ysb, int lane) {
    const bool scatter = p.off_class >= 0 || p.off_roff >= 0 || p.off_box >= 0 || p.off_track >= 0 || p.extra_n > 0;
    for (int j = o_begin + kMaxObjSmem; j < o_end; ++j) {
        ObjDerived d;
        derive(p.objs[j], p, d);
        d.tab = d.tabr = -1;
        if (d.x0 < d.x1 && max(d.y0, ysa) < min(d.y1, ysb + 1)) splat_object_slow(p, d, nullptr, nullptr, st, q0, s0, s1, ysa, ysb, lane);
        const int qs = d.scy * p.W + d.scx;
        if (scatter && d.scx >= 0 && qs >= s0 && qs < s1) {   // warp-uniform: the warp that owns the centre pixel
            bool later = false;
            for (int j2 = j + 1 + lane; j2 < o_end; j2 += 32) {
                ObjDerived e;
                derive(p.objs[j2], p, e);
                later |= e.scx == d.scx && e.scy == d.scy;
            }
            d.last_at_pixel = !__any_sync(0xffffffffu, later);
            if (lane == 0) scatter_object(p, d, j, st, q0, q1);
        }
    }
}

// ---- a builder group: fill -> splat -> scatter -> ignore areas -> one bulk store, chunk after chunk ----
__device__ __forceinline__ void builder_main(const RenderParams& p, RenderShared& S, float* st, int g, int gt, long long c0,
                                             long long c1, int img0) {
    const int lane = gt & 31, wg = gt >> 5;
    const int W = p.W, HW = p.HW, Cout = p.Cout, wch = p.wch, P = p.P, cpi = p.cpi;
    const bool has_w = wch >= 0;
    const int bar_id = 1 + g;
    // pattern fill: thread t < n_fill writes the float4s t, t + n_fill, ...; n_fill is a multiple of Cout, so the channel
    // phase of its float4 -- and with it the value (zeros, 1.0 where the weights channel falls) -- never changes
    const int n_fill = (kGT / Cout) * Cout;
    float4 fill_v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (has_w) {
        const int r = (4 * gt) % Cout;   // channel of the first float of this thread's float4s
        fill_v.x = ((r + 0) % Cout == wch) ? 1.f : 0.f;
        fill_v.y = ((r + 1) % Cout == wch) ? 1.f : 0.f;
        fill_v.z = ((r + 2) % Cout == wch) ? 1.f : 0.f;
        fill_v.w = ((r + 3) % Cout == wch) ? 1.f : 0.f;
    }
    const bool scatter = p.off_class >= 0 || p.off_roff >= 0 || p.off_box >= 0 || p.off_track >= 0 || p.extra_n > 0;
    int cur_rel = -1;
#ifdef CVM_EXPERIMENT
    long long dbg_t[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}, dbg_last = clock64();
#endif
    // (image, chunk in image) of the group's first chunk; advanced by kGroups chunks per iteration without divisions
    int img = (int)((c0 + g) / cpi), ci = (int)((c0 + g) - (long long)img * cpi);
    for (long long c = c0 + g; c < c1; c += kGroups) {
        const int rel = img - img0;
        const int q0 = ci * P, q1 = min(HW, q0 + P), npx = q1 - q0;
        const int ya = q0 / W, xa = q0 - ya * W;   // row / column of the chunk's first pixel
        int yb = ya;                               // row of its last pixel
        for (int t = xa + npx - 1; t >= W; t -= W) ++yb;
        if (gt == 0) {
            if (rel != cur_rel) {   // (every reader of the previous image's set passed the last barrier of its last chunk)
                fence_cta_r();
                S.prog[g] = rel;
            }
            if (p.bulk) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // the previous chunk has left the buffer
        }
        RCLK(0);
        bar_group(bar_id, kGT);   // ---- barrier C: the buffer is free ----
        RCLK(1);
        // ---- phase 1: pattern fill (heat = 0, weights = 1, regression targets = 0; processor.py:267-268) ----
        if (Cout > kGT) {   // (very wide pixels: no thread keeps a constant channel phase; scalar fill)
            for (int f = gt; f < npx * Cout; f += kGT) st[f] = (has_w && f % Cout == wch) ? 1.f : 0.f;
        } else if (gt < n_fill && !RDBG(4)) {
            const int n4 = (npx * Cout + 3) >> 2;   // the buffer is a whole number of float4s
            float4* s4 = reinterpret_cast<float4*>(st);
#pragma unroll 4
            for (int f = gt; f < n4; f += n_fill) s4[f] = fill_v;
        }
        RCLK(2);
        if (rel != cur_rel) {
            if (gt == 0) {   // the image's tables (built by the setup group while this group filled)
                RENDER_SPIN(S.ready > rel);
                fence_cta_r();
            }
            cur_rel = rel;
        }
        RCLK(3);
        bar_group(bar_id, kGT);   // ---- barrier A: the fill is complete, the table set is visible ----
        RCLK(4);
        const TableSet& T = S.set[rel & 1];
        const int n = T.n, o_begin = T.o_begin, o_end = T.o_end, n_ign = T.n_ign, i_begin = T.i_begin;

        // ---- phase 2: every warp owns a contiguous quarter of the chunk's pixels and max/min-combines into it every
        //      object that meets its rows (found with a ballot, lane = object), one after the other ----
        if (!RDBG(1)) {
            const int per = (npx + kGW - 1) / kGW;
            const int s0 = q0 + wg * per, s1 = min(q1, s0 + per);
            int ysa = ya, ysb;   // rows of the first and last pixel of the piece
            int t = xa + wg * per;
            for (; t >= W; t -= W) ++ysa;
            ysb = ysa;
            for (t += s1 - s0 - 1; t >= W; t -= W) ++ysb;
            if (s0 < s1) {
                for (int b0 = 0; b0 < n; b0 += 32) {
                    const int o = b0 + lane;
                    bool meets = false;
                    if (o < n) {   // does the window meet this warp's pixels?  (row by row: the piece may start / end inside a row)
                        const int4 w = T.win[o];
                        for (int y = max(w.z, ysa), ye = min(w.w, ysb + 1); y < ye && !meets; ++y)
                            meets = max(w.x, s0 - y * W) < min(w.y, s1 - y * W);
                    }
                    unsigned m = __ballot_sync(0xffffffffu, meets);
#ifdef CVM_EXPERIMENT
                    RCLK(8);
                    dbg_t[9] += __popc(m);
#endif
                    while (m) {
                        const int b = __ffs(m) - 1;
                        m &= m - 1;
                        splat_object(p, T.obj[b0 + b], T.fac, T.fac, st, q0, s0, s1, ysa, ysb, lane);
                    }
                }
            }
            RCLK(10);
            if (scatter && gt < n) scatter_object(p, T.obj[gt], o_begin + gt, st, q0, q1);
            if (o_end - o_begin > kMaxObjSmem && s0 < s1) splat_extra_objects(p, st, o_begin, o_end, q0, q1, s0, s1, ysa, ysb, lane);
        }

        RCLK(5);
        // ---- phase 3: ignore areas: weights = 0, input-px numbers used as mask indices (processor.py:318-323) ----
        bool ign_hit = n_ign > kMaxIgnSmem;   // (boxes beyond the cache: take the slow way)
        for (int i = 0; i < min(n_ign, kMaxIgnSmem) && !ign_hit; ++i)
            ign_hit = T.ign[i][0] < T.ign[i][1] && max(T.ign[i][2], ya) < min(T.ign[i][3], yb + 1);
        if (ign_hit) {   // uniform in the group: few chunks meet an ignore box
            bar_group(bar_id, kGT);   // all splats are in
            for (int i = 0; i < n_ign; ++i) {
                int sx, ex, sy, ey;
                if (i < kMaxIgnSmem) {
                    sx = T.ign[i][0], ex = T.ign[i][1], sy = T.ign[i][2], ey = T.ign[i][3];
                } else {
                    const cvm_box bx = p.ignore[i_begin + i];
                    sx = max((int)bx.x, 0), ex = min(max((int)(bx.x + bx.w), 0), W);
                    sy = max((int)bx.y, 0), ey = min(max((int)(bx.y + bx.h), 0), p.H);
                }
                sy = max(sy, ya);
                ey = min(ey, yb + 1);
                const int bw = ex - sx, cells = bw * (ey - sy);
                if (bw <= 0 || ey <= sy) continue;
                for (int k = gt; k < cells; k += kGT) {
                    const int yy = sy + k / bw, xx = sx + k % bw, q = yy * W + xx;
                    if (q >= q0 && q < q1) st[(size_t)(q - q0) * Cout + wch] = 0.f;
                }
            }
        }

        // ---- stream the chunk out ----
        float* const dst = p.out + ((size_t)img * HW + q0) * Cout;
        if (p.bulk) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the bulk engine
            RCLK(6);
            bar_group(bar_id, kGT);   // ---- barrier B: the chunk is complete ----
            RCLK(7);
            if (gt == 0 && !RDBG(2)) bulk_s2g(dst, st, (uint32_t)(npx * Cout * 4));
        } else {
            bar_group(bar_id, kGT);
            const int nf = npx * Cout;
            for (int f = gt; f < nf; f += kGT) dst[f] = st[f];   // (barrier C of the next chunk: the buffer has been read)
        }
        ci += kGroups;
        while (ci >= cpi) {
            ci -= cpi;
            ++img;
        }
    }
#ifdef CVM_EXPERIMENT
    if (gt == 0)
        for (int i = 0; i < 12; ++i) atomicAdd(&g_render_dbg[i], (unsigned long long)dbg_t[i]);
#endif
    if (gt == 0) {
        fence_cta_r();
        S.prog[g] = 0x7fffffff;   // this group reads no table set any more
        if (p.bulk) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // stores complete before the CTA retires
    }
}

// One persistent CTA per SM owns a contiguous range of chunks (P consecutive pixels of one image, all channels).  Its
// builder groups take the chunks of the range round-robin and never synchronise with each other (named barriers per
// group), so while one group's bulk store drains, the others fill and splat: the store engine of the SM always has a chunk
// queued.  The setup group prepares the per-image tables one image ahead.
__global__ void __launch_bounds__(kThreads, 1) render_kernel(const __grid_constant__ RenderParams p) {
    RenderShared& S = *reinterpret_cast<RenderShared*>(g_render_smem);
    constexpr size_t kStageOff = (sizeof(RenderShared) + 127) & ~(size_t)127;
    const int tid = threadIdx.x;
    const long long G = gridDim.x, cta = blockIdx.x;
    const long long c0 = cta * p.n_chunks / G, c1 = (cta + 1) * p.n_chunks / G;
    if (tid < kGroups) S.prog[tid] = 0;
    if (c0 >= c1) return;
    const int img0 = (int)(c0 / p.cpi), n_img = (int)((c1 - 1) / p.cpi) - img0 + 1;
    setup_image(p, S, S.set[0], img0, tid, kThreads, 0);   // the first image: all threads, nobody has anything else to do yet
    __syncthreads();
    if (tid == 0) S.ready = 1;
    const int g = tid / kGT, gt = tid - g * kGT;
    if (g >= kGroups) setup_main(p, S, tid - kGroups * kGT, img0, n_img);
    else builder_main(p, S, reinterpret_cast<float*>(g_render_smem + kStageOff + (size_t)g * kBufBytes), g, gt, c0, c1, img0);
}

int launch_render(const RenderParams& p0, int B, cudaStream_t st) {
    RenderParams p = p0;
    p.HW = p.H * p.W;
    CVM_CHECK_ARG(p.Cout * 16 <= kBufBytes, "render: %d output channels do not fit the staging buffer", p.Cout);
    // chunks of (nearly) equal size: as few as fit the buffer, then the pixels of an image spread evenly over them
    const int hw4 = (p.HW + 3) & ~3;
    int pmax = (kBufBytes / (p.Cout * 4)) & ~3;
    if (pmax > hw4) pmax = hw4;
    const int cpi0 = (p.HW + pmax - 1) / pmax;
    int P = ((p.HW + cpi0 - 1) / cpi0 + 3) & ~3;
    if (P > pmax) P = pmax;
    p.P = P;
    p.cpi = (p.HW + P - 1) / P;
    p.n_chunks = (long long)B * p.cpi;
    if (p.n_chunks == 0) return CVM_OK;
    // bulk stores need 16-byte granules: base aligned, every image a whole number of them (chunks are P*Cout*4 bytes with
    // P % 4 == 0, the partial last chunk of an image then ends on a granule too)
    p.bulk = cvm_aligned16(p.out) && (((long long)p.HW * p.Cout) % 4 == 0);
#ifdef CVM_EXPERIMENT
    if (const char* e = getenv("CVM_RENDER_SKIP")) p.dbg_skip = atoi(e);
    if RDBG(8) p.bulk = 0;   // experiment: plain stores instead of bulk copies
#endif
    const size_t smem = ((sizeof(RenderShared) + 127) & ~(size_t)127) + (size_t)kGroups * kBufBytes;
    CVM_SMEM_ATTR_ONCE((render_kernel), smem);
    long long grid = cvm_num_sms();
    if (grid > p.n_chunks) grid = p.n_chunks;
    render_kernel<<<(unsigned)grid, kThreads, smem, st>>>(p);
    CVM_CHECK_LAUNCH("render_kernel");
    return CVM_OK;
}

}  // namespace

#ifdef CVM_EXPERIMENT
extern "C" int cvm_render_debug(unsigned long long* out8) {   // reads and clears the phase counters
    unsigned long long z[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    cudaMemcpyFromSymbol(out8, g_render_dbg, sizeof(z));
    cudaMemcpyToSymbol(g_render_dbg, z, sizeof(z));
    return 0;
}
#endif

extern "C" int cvm_render_gt(const cvm_layout* L, const cvm_obj* objs, const int32_t* obj_offsets, const cvm_box* ignore,
                             const int32_t* ign_offsets, int B, float* y_true, void* stream) {
    return cvm_render_gt_extra(L, objs, obj_offsets, ignore, ign_offsets, B, nullptr, 0, 0, 0, y_true, stream);
}

extern "C" int cvm_render_gt_extra(const cvm_layout* L, const cvm_obj* objs, const int32_t* obj_offsets, const cvm_box* ignore,
                                   const int32_t* ign_offsets, int B, const float* extra, int extra_stride, int extra_off,
                                   int extra_n, float* y_true, void* stream) {
    CVM_CHECK_ARG(L && obj_offsets && y_true, "NULL pointer argument");
    CVM_CHECK_ARG(extra_n == 0 || (extra && extra_n > 0 && extra_stride >= extra_n && extra_off >= L->hm && extra_off + extra_n <= L->Cp),
                  "extra targets [%d, %d) must lie inside the regression channels [hm, Cp)", extra_off, extra_off + extra_n);
    CVM_CHECK_ARG(B >= 0 && L->H > 0 && L->W > 0, "bad shape");
    CVM_CHECK_ARG(L->hm >= 1 && L->hm <= 64 && L->Ct == L->Cp + 1 && L->Cp >= L->hm, "bad channel layout");
    CVM_CHECK_ARG((ignore == nullptr) == (ign_offsets == nullptr), "ignore and ign_offsets must both be given or both NULL");
    RenderParams p;
    memset(&p, 0, sizeof(p));
    p.objs = objs;
    p.obj_offsets = obj_offsets;
    p.ignore = ignore;
    p.ign_offsets = ign_offsets;
    p.out = y_true;
    p.H = L->H;
    p.W = L->W;
    p.Cout = L->Ct;
    p.n_planes = L->hm;
    p.wch = L->Ct - 1;
    p.per_class = L->hm > 1;
    p.force_explicit = 0;
    p.off_class = L->off_class;
    p.off_roff = L->off_roff;
    p.off_box = L->off_box;
    p.off_track = L->off_track;
    p.extra = extra_n > 0 ? extra : nullptr;
    p.extra_stride = extra_stride;
    p.extra_off = extra_off;
    p.extra_n = extra_n;
    p.R = L->R;
    p.alpha = L->alpha;
    return launch_render(p, B, reinterpret_cast<cudaStream_t>(stream));
}

extern "C" int cvm_render_prev_hm(const cvm_layout* L, const cvm_obj* objs, const int32_t* obj_offsets, int B,
                                  float* prev_hm, void* stream) {
    CVM_CHECK_ARG(L && obj_offsets && prev_hm, "NULL pointer argument");
    CVM_CHECK_ARG(B >= 0 && L->H > 0 && L->W > 0, "bad shape");
    RenderParams p;
    memset(&p, 0, sizeof(p));
    p.objs = objs;
    p.obj_offsets = obj_offsets;
    p.out = prev_hm;
    p.H = L->H;
    p.W = L->W;
    p.Cout = 1;
    p.n_planes = 1;
    p.wch = -1;
    p.per_class = 0;
    p.force_explicit = 1;
    p.off_class = p.off_roff = p.off_box = p.off_track = -1;
    p.R = L->R;
    p.alpha = L->alpha;
    return launch_render(p, B, reinterpret_cast<cudaStream_t>(stream));
}

// =====================================================================================================================
// In-place variant for the literal fill_heatmap drop-in (processor.py:18-22 mutates caller-owned arrays): the objects
// are max/min-combined INTO existing planes.  One thread per pixel walks the object list, so no two threads ever touch
// the same cell.  Not the fast path (that is cvm_render_gt); it exists so single calls keep the reference semantics.
// =====================================================================================================================
namespace {

__global__ void __launch_bounds__(256) fill_inplace_kernel(const cvm_obj* __restrict__ objs, int n_obj, float* heat,
                                                           int heat_stride, float* weights, int H, int W, double R,
                                                           double alpha) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= H * W) return;
    const int y = i / W, x = i - y * W;
    float hv = heat[(size_t)i * heat_stride];
    float wv = weights ? weights[i] : 1.f;
    for (int o = 0; o < n_obj; ++o) {
        const cvm_obj ob = objs[o];
        const double w = ob.w, h = ob.h;
        const int hx = (int)floor(w / 2.0), hy = (int)floor(h / 2.0);
        if (x < max(0, ob.cx - hx) || x >= min(W, ob.cx + hx) || y < max(0, ob.cy - hy) || y >= min(H, ob.cy + hy)) continue;
        const double sdx = (alpha * w) / (6.0 * R), sdy = (alpha * h) / (6.0 * R);
        const double dx = (double)(x - ob.cx), dy = (double)(y - ob.cy);
        const double g = exp(-(dx * dx / (2.0 * (sdx * sdx)) + dy * dy / (2.0 * (sdy * sdy))));
        const double rw = 1.0 - ((fmin(20.0, fmax(w, h)) / 16.0) - 0.25);
        hv = fmaxf(hv, (float)(g * (double)ob.peak));
        wv = fminf(wv, (float)(1.0 - rw * g));
    }
    heat[(size_t)i * heat_stride] = hv;
    if (weights) weights[i] = wv;
}

}  // namespace

extern "C" int cvm_fill_heatmap_inplace(const cvm_obj* objs, int n_obj, float* heat, int heat_stride, float* weights,
                                        int H, int W, double R, double alpha, void* stream) {
    CVM_CHECK_ARG(objs && heat && n_obj >= 0 && H > 0 && W > 0 && heat_stride >= 1, "bad argument");
    if (n_obj == 0) return CVM_OK;
    fill_inplace_kernel<<<(H * W + 255) / 256, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        objs, n_obj, heat, heat_stride, weights, H, W, R, alpha);
    CVM_CHECK_LAUNCH("fill_inplace_kernel");
    return CVM_OK;
}
