// Ground-truth render: gaussian splat (max-combine) + loss-weights (min-combine) + centre scatter + ignore areas.
//
// Replaces fill_heatmap (reference models/centernet/processor.py:17-38, a numba double loop per object) and the render
// part of ProcessImages.process (processor.py:264-334) for a whole batch; also gen_prev_heatmap
// (models/centertracker/processor.py:22-41) with one plane and no weights plane.
//
// Bound: HBM writes.  Algorithmic bytes per image: 4*H*W*Ct, every byte written exactly once (zeros included).
//
// Structure: y_true is one flat list of chunks (P consecutive pixels of one image, all channels: a contiguous piece of
// the NHWC tensor).  The list is cut into equal contiguous ranges, one persistent CTA per range (two per SM).  A chunk is
// built in shared memory in its FINAL layout: a pattern fill (zeros, ones in the weights channel); then one warp per
// (object, 32-column segment) work unit walks the unit's rows, lane = column, and max-/min-combines the gaussian into the
// chunk with shared integer atomics on the float bit patterns (windows of different objects overlap); then ignore areas
// and the handful of regression targets at centre pixels are written, and ONE bulk async copy (TMA engine: UBLKCP,
// shared -> global) streams the chunk out while the SM's other CTA builds its chunk.  No store instruction
// touches global memory on the fast path.  All gaussian math is fp64 like the reference's (scalar fp64 stored as fp32);
// the separable factors exp(-ax), exp(-ay) are tabulated (per image / per chunk), see render_kernel.
// (Tried in round 2: the per-chunk integer bookkeeping - ranges, first / last row, ignore-box test, a fifth of the kernel's
// instructions because all 32 warps of an SM repeat it - done by ONE thread a chunk ahead and read from shared memory:
// 0.173 ms instead of 0.163; the kernel is bound by the latency of its barrier-separated phases, and the lone thread
// lengthens exactly that.  Also tried: more, smaller CTAs (3 x 384 threads with 44 KB chunks: 0.168 ms; 2 x 256 / 320 / 384 /
// 448 / 512 threads: 0.173 / 0.168 / 0.158 / 0.158 / 0.159) and a pixel-parallel splat in which every thread owns pixels of
// the chunk and walks the objects that meet its rows with plain max / min instead of warp-per-unit shared atomics: bit-exact
// too, but 0.21-0.28 ms - 15 k window tests per chunk against 1.3 k covered cells, and the fp64 path does not fit 64
// registers.)
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
constexpr int kThreads = 512;
constexpr int kWarps = kThreads / 32;
constexpr int kMaxObjSmem = 64;   // objects are processed in batches of this many
constexpr int kMaxIgnSmem = 16;     // ignore boxes cached in shared memory per image (more are read from global)
#ifndef CVM_RENDER_CHUNK
#define CVM_RENDER_CHUNK 81920
#endif
#ifndef CVM_RENDER_COLTAB
#define CVM_RENDER_COLTAB 1536   /* the 32 objects of BASELINE configs[1] need ~1350 column entries; 80 KB chunks make room: 0.1606 -> 0.1571 ms */
#endif
constexpr int kChunkBytes = CVM_RENDER_CHUNK;   // staging buffer per CTA (two CTAs per SM: one builds while the other's chunk streams out)
constexpr int kMaxUnits = 512;       // (object, 32-column segment) work units per chunk and object batch
constexpr int kColTab = CVM_RENDER_COLTAB;        // entries of the per-image column-factor table (objects that do not fit use exp)
constexpr int kRowTab = 1024;        // entries of the per-image row-factor table

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
}

// shared -> global bulk copy (bytes % 16 == 0, both addresses 16-byte aligned), tracked by the issuing thread's bulk groups
__device__ __forceinline__ void bulk_s2g(void* dst_gmem, const void* src_smem, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem), "r"(smem_u32(src_smem)), "r"(bytes)
                 : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}

extern __shared__ __align__(128) unsigned char g_render_smem[];   // the staging buffer (kChunkBytes)

// max / min combine of a float into shared memory through integer atomics on the bit pattern: non-negative floats order
// like signed ints, negative floats order inversely like unsigned ints, and each of the two operations keeps the cell
// monotone in float order, so any interleaving of them ends at the true max / min.
__device__ __forceinline__ void smem_max_float(float* cell, float v) {
    if (v >= 0.f) atomicMax(reinterpret_cast<int*>(cell), __float_as_int(v));
    else atomicMin(reinterpret_cast<unsigned int*>(cell), __float_as_uint(v));
}
__device__ __forceinline__ void smem_min_float(float* cell, float v) {
    if (v >= 0.f) atomicMin(reinterpret_cast<int*>(cell), __float_as_int(v));
    else atomicMax(reinterpret_cast<unsigned int*>(cell), __float_as_uint(v));
}

// The gaussian is separable: exp(-(ax + ay)) = exp(-ax) * exp(-ay).  The column factors exp(-ax) of every object are
// and its row factors exp(-ay) are tabulated once per image, so a covered pixel costs one fp64 multiply instead
// of one fp64 exp.  The product differs from the exp of the sum by a few ulp of a DOUBLE; after the rounding to fp32
// that the reference stores, the value is the same except when a rounding boundary falls inside that interval
// (probability ~1e-8 per value), where it differs by one fp32 ulp -- far inside the 1e-5 tolerance.
__global__ void __launch_bounds__(kThreads, 2) render_kernel(const __grid_constant__ RenderParams p) {
    __shared__ ObjDerived sobj[kMaxObjSmem];
    __shared__ double s_col[kColTab];                   // column factors, object after object
    __shared__ double s_row[kRowTab];                   // row factors, object after object
    __shared__ int s_unit[kMaxUnits];                   // work units: object (low 8 bits) | 32-column segment of its window
    __shared__ int s_nunits[2];                         // double-buffered by list-build parity (reset one build ahead)
    __shared__ int s_ign[kMaxIgnSmem][4];               // ignore boxes of the image: sx, ex, sy, ey (clamped to the map)
    __shared__ unsigned char s_own[kColTab + kRowTab];   // object that owns each table entry
    __shared__ int s_used[2];                            // table entries in use (columns, rows)

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int W = p.W, HW = p.HW, Cout = p.Cout, wch = p.wch, P = p.P, cpi = p.cpi;
    const bool has_w = wch >= 0;
    float* const stage0 = reinterpret_cast<float*>(g_render_smem);

    const long long G = gridDim.x, g = blockIdx.x;
    const long long c0 = g * p.n_chunks / G, c1 = (g + 1) * p.n_chunks / G;
    // pattern fill: thread t < n_fill writes the float4s t, t + n_fill, ...; n_fill is a multiple of Cout, so the channel
    // phase of its float4 -- and with it the value (zeros, 1.0 where the weights channel falls) -- never changes
    const int n_fill = (kThreads / Cout) * Cout;
    float4 fill_v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (has_w) {
        const int r = (4 * tid) % Cout;   // channel of the first float of this thread's float4s
        fill_v.x = ((r + 0) % Cout == wch) ? 1.f : 0.f;
        fill_v.y = ((r + 1) % Cout == wch) ? 1.f : 0.f;
        fill_v.z = ((r + 2) % Cout == wch) ? 1.f : 0.f;
        fill_v.w = ((r + 3) % Cout == wch) ? 1.f : 0.f;
    }
    const bool scatter = p.off_class >= 0 || p.off_roff >= 0 || p.off_box >= 0 || p.off_track >= 0 || p.extra_n > 0;
    int loaded_img = -1;         // image whose per-image state (ranges, ignore boxes, first object batch) is loaded
    int loaded_base = -1;        // first object of the batch in sobj
    int o_begin = 0, o_end = 0, i_begin = 0, n_ign = 0, n_batches = 0;
    int lk = 0;                  // list builds so far (parity picks the counter)
    const int step_rows = P / W, step_cols = P - step_rows * W;   // how (ya, xa) advance from one chunk to the next
    if (tid < 2) s_nunits[tid] = 0;
    __syncthreads();

    int img = (int)(c0 / cpi);
    int ci = (int)(c0 - (long long)img * cpi);
    int ya = (ci * P) / W, xa = ci * P - ya * W;   // row / column of the chunk's first pixel, kept incrementally
    const int n_local = (int)(c1 - c0);
    for (int it = 0; it < n_local; ++it) {
        const int q0 = ci * P, q1 = min(HW, q0 + P), npx = q1 - q0;
        int yb = ya;   // row of the chunk's last pixel
        for (int t = xa + npx - 1; t >= W; t -= W) ++yb;
        float* const st = stage0;
        const bool new_img = loaded_img != img;
        if (new_img) {   // object / ignore-box ranges of the image: fetched once per image, not once per chunk
            o_begin = p.obj_offsets[img];
            o_end = p.obj_offsets[img + 1];
            n_batches = (o_end - o_begin + kMaxObjSmem - 1) / kMaxObjSmem;
            n_ign = i_begin = 0;
            if (p.ignore != nullptr && has_w) {
                i_begin = p.ign_offsets[img];
                n_ign = p.ign_offsets[img + 1] - i_begin;
            }
        }

        if (new_img) {   // clamp the ignore boxes once per image (read after the barriers below)
            if (tid < min(n_ign, kMaxIgnSmem)) {
                const cvm_box bx = p.ignore[i_begin + tid];
                s_ign[tid][0] = max((int)bx.x, 0);
                s_ign[tid][1] = min(max((int)(bx.x + bx.w), 0), W);
                s_ign[tid][2] = max((int)bx.y, 0);
                s_ign[tid][3] = min(max((int)(bx.y + bx.h), 0), p.H);
            }
        }


        // The pre-work of a chunk - per-image tables and the chunk's work units, nothing that touches the staging buffer -
        // runs BEFORE the wait for the buffer, i.e. while the bulk store of the previous chunk is still reading it (every
        // reader of these arrays finished before barrier B of the previous chunk).  Batches after the first one of a
        // crowded image (rare) do the same work inline.
        for (int bi = 0;; ++bi) {
            const bool have = bi < n_batches;
            const int base = o_begin + bi * kMaxObjSmem;
            const int n = have ? min(kMaxObjSmem, o_end - base) : 0;
            const int par = have ? (lk++) & 1 : 0;
            if (have) {
                if (new_img || loaded_base != base) {
                    // ---- once per image (and object batch): derived records, scatter winners, column factors ----
                    if (bi > 0) __syncthreads();   // the previous batch is still being read
                    if (tid < n) derive(p.objs[base + tid], p, sobj[tid]);
                    __syncthreads();
                    if (tid < n && scatter) {
                        // a later object (list order) on the same centre pixel overwrites r_offset / fullbox / track_offset
                        // (processor.py:288-299), so only the last one writes them
                        const int sx = sobj[tid].scx, sy = sobj[tid].scy;
                        bool last = true;
                        for (int j = base + tid + 1; j < o_end && last; ++j) {
                            if (j - base < n) {
                                if (sobj[j - base].scx == sx && sobj[j - base].scy == sy) last = false;
                            } else {
                                ObjDerived e;
                                derive(p.objs[j], p, e);
                                if (e.scx == sx && e.scy == sy) last = false;
                            }
                        }
                        sobj[tid].last_at_pixel = last;
                    }
                    if (tid < n) {   // table space in object order; an object that does not fit evaluates exp per pixel
                        int off = 0, offr = 0;
                        for (int o = 0; o < tid; ++o) {
                            off += max(0, sobj[o].x1 - sobj[o].x0);
                            offr += max(0, sobj[o].y1 - sobj[o].y0);
                        }
                        const int wd = max(0, sobj[tid].x1 - sobj[tid].x0), ht = max(0, sobj[tid].y1 - sobj[tid].y0);
                        const int tab = off + wd <= kColTab ? off : -1, tabr = offr + ht <= kRowTab ? offr : -1;
                        // owner of every table entry, so that the factors can be computed one entry per thread
                        if (tab >= 0)
                            for (int k = 0; k < wd; ++k) s_own[tab + k] = (unsigned char)tid;
                        if (tabr >= 0)
                            for (int k = 0; k < ht; ++k) s_own[kColTab + tabr + k] = (unsigned char)tid;
                        if (tid == n - 1) {
                            s_used[0] = tab >= 0 ? tab + wd : 0;      // (the last object fits only if all before it did)
                            s_used[1] = tabr >= 0 ? tabr + ht : 0;
                        }
                        sobj[tid].tab = tab;
                        sobj[tid].tabr = tabr;
                    }
                    __syncthreads();
                    if (sobj[n - 1].tab < 0 || sobj[n - 1].tabr < 0) {
                        // some object did not fit: its entries are not tabulated, the entries before it are found by scanning
                        if (tid == 0) {
                            int used = 0, usedr = 0;
                            for (int o = 0; o < n; ++o) {
                                if (sobj[o].tab >= 0) used = sobj[o].tab + max(0, sobj[o].x1 - sobj[o].x0);
                                if (sobj[o].tabr >= 0) usedr = sobj[o].tabr + max(0, sobj[o].y1 - sobj[o].y0);
                            }
                            s_used[0] = used;
                            s_used[1] = usedr;
                        }
                        __syncthreads();
                    }
                    {
                        const int used = s_used[0], usedr = s_used[1];
                        for (int e = tid; e < used + usedr; e += kThreads) {
                            if (e < used) {
                                const ObjDerived& d = sobj[s_own[e]];
                                const double dx = (double)(d.x0 + (e - d.tab) - d.cx);
                                s_col[e] = exp(-(dx * dx * d.inv2vx));
                            } else {
                                const int r = e - used;
                                const ObjDerived& d = sobj[s_own[kColTab + r]];
                                const double dy = (double)(d.y0 + (r - d.tabr) - d.cy);
                                s_row[r] = exp(-(dy * dy * d.inv2vy));
                            }
                        }
                    }
                    loaded_base = base;
                }
                // ---- work units of this chunk ----
                if (tid < n && !RDBG(32)) {
                    const ObjDerived& d = sobj[tid];
                    if (max(d.y0, ya) < min(d.y1, yb + 1) && d.x0 < d.x1) {
                        const int u = (d.x1 - d.x0 + 31) >> 5;
                        const int start = atomicAdd(&s_nunits[par], u);
                        for (int k = 0; k < u && start + k < kMaxUnits; ++k) s_unit[start + k] = tid | (k << 8);
                    }
                }
                if (tid == 0) s_nunits[par ^ 1] = 0;   // the other counter is idle: reset it for the next build
            }
            if (bi == 0) {
                if (p.bulk && tid == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // the previous chunk has left the buffer
                __syncthreads();   // ---- barrier C: the buffer is free ----
                // ---- phase 1: pattern fill (heat = 0, weights = 1, regression targets = 0; processor.py:267-268) ----
                if (tid < n_fill && !RDBG(4)) {
                    const int n4 = (npx * Cout + 3) >> 2;   // the buffer is a whole number of float4s
                    float4* s4 = reinterpret_cast<float4*>(st);
                    for (int f = tid; f < n4; f += n_fill) s4[f] = fill_v;
                }
            }
            if (!have) break;
            __syncthreads();   // ---- barrier A: fill, tables and units are visible ----

            // ---- phase 2: one warp per (unit, row) item, lane = column; max/min-combine with shared atomics (windows of
            //      different objects overlap).  Items are dealt round-robin to the warps; order is irrelevant ----
            const int total = s_nunits[par];
            const bool overflow = total > kMaxUnits;   // absurdly many wide objects: whole-object items instead
            const int n_items = RDBG(1) ? 0 : (overflow ? n : total);
            for (int it = warp; it < n_items; it += kWarps) {
                const int u = overflow ? it : s_unit[it];
                const ObjDerived& d = sobj[u & 255];
                // the record is read ONCE: the shared atomics below are memory clobbers, anything read through `d` after
                // them would be reloaded row after row (a chain of dependent shared loads per row)
                const int x0 = d.x0, x1 = d.x1, y0 = d.y0, tab = d.tab, tabr = d.tabr;
                const int r0 = max(y0, ya), r1 = min(d.y1, yb + 1);
                float* const plane_px = st + d.plane;
                const double peak = d.peak, rw = d.rw;
                if (x0 >= x1) continue;
                int xs = x0, xe = x1;
                if (!overflow) {
                    xs = x0 + ((u >> 8) << 5);
                    xe = min(x1, xs + 32);
                }
                for (int x = xs + lane; x < xe; x += 32) {
                    double ex;
                    if (tab >= 0) {
                        ex = s_col[tab + (x - x0)];
                    } else {
                        const double dx = (double)(x - d.cx);
                        ex = exp(-(dx * dx * d.inv2vx));
                    }
                    for (int y = r0; y < r1; ++y) {
                        const int q = y * W + x;
                        if (q < q0 || q >= q1) continue;   // the chunk may start / end inside a row
                        double ey;
                        if (tabr >= 0) {
                            ey = s_row[tabr + (y - y0)];
                        } else {
                            const double dy = (double)(y - d.cy);
                            ey = exp(-(dy * dy * d.inv2vy));
                        }
                        const double gv = ex * ey;                                           // processor.py:34-36
                        const int o = (q - q0) * Cout;
                        smem_max_float(plane_px + o, (float)(gv * peak));                     // :37
                        if (has_w) smem_min_float(st + o + wch, (float)(1.0 - rw * gv));      // :38
                    }
                }
            }
            // centre scatter: regression targets and the class one-hot live in channels the splat never touches
            if (scatter && tid < n && !RDBG(64)) {
                const ObjDerived& d = sobj[tid];
                const int qs = d.scy * W + d.scx;
                if (d.scx >= 0 && qs >= q0 && qs < q1) {
                    float* px = st + (size_t)(qs - q0) * Cout;
                    if (p.off_class >= 0 && d.cls >= 0 && p.off_class + d.cls < Cout) px[p.off_class + d.cls] = 1.0f;   // :290
                    if (d.last_at_pixel) {
                        if (p.off_roff >= 0) {
                            px[p.off_roff] = d.offx;                                                                   // :292
                            px[p.off_roff + 1] = d.offy;
                        }
                        if (p.off_box >= 0) {
                            px[p.off_box] = d.bw;                                                                      // :294
                            px[p.off_box + 1] = d.bh;
                        }
                        if (p.off_track >= 0) {
                            px[p.off_track] = d.tx;
                            px[p.off_track + 1] = d.ty;
                        }
                        if (p.extra_n > 0) {   // l_shape (7) / 3d_info (5) targets, computed on the host (processor.py:69-115,296-299)
                            const float* e = p.extra + (size_t)(base + tid) * p.extra_stride;
                            for (int k = 0; k < p.extra_n; ++k) px[p.extra_off + k] = e[k];
                        }
                    }
                }
            }
        }

        // ---- phase 3: ignore areas: weights = 0, input-px numbers used as mask indices (processor.py:318-323) ----
        bool ign_hit = n_ign > kMaxIgnSmem || (n_ign > 0 && n_batches == 0);   // (boxes beyond the cache / no barrier yet: take the slow way)
        for (int i = 0; i < (RDBG(128) ? 0 : min(n_ign, kMaxIgnSmem)) && !ign_hit; ++i)
            ign_hit = s_ign[i][0] < s_ign[i][1] && max(s_ign[i][2], ya) < min(s_ign[i][3], yb + 1);
        if (ign_hit) {   // uniform: few chunks meet an ignore box
            __syncthreads();   // all splats are in (and s_ign is visible when the image had no objects)
            for (int i = 0; i < n_ign; ++i) {
                int sx, ex, sy, ey;
                if (i < kMaxIgnSmem) {
                    sx = s_ign[i][0], ex = s_ign[i][1], sy = s_ign[i][2], ey = s_ign[i][3];
                } else {
                    const cvm_box bx = p.ignore[i_begin + i];
                    sx = max((int)bx.x, 0), ex = min(max((int)(bx.x + bx.w), 0), W);
                    sy = max((int)bx.y, 0), ey = min(max((int)(bx.y + bx.h), 0), p.H);
                }
                sy = max(sy, ya);
                ey = min(ey, yb + 1);
                const int bw = ex - sx, cells = bw * (ey - sy);
                if (bw <= 0 || ey <= sy) continue;
                for (int k = tid; k < cells; k += kThreads) {
                    const int yy = sy + k / bw, xx = sx + k % bw, q = yy * W + xx;
                    if (q >= q0 && q < q1) st[(size_t)(q - q0) * Cout + wch] = 0.f;
                }
            }
        }

        // ---- stream the chunk out ----
        if (p.bulk) {
            if (!RDBG(16)) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the bulk engine
            __syncthreads();   // ---- barrier B: the chunk is complete ----
            // the buffer is rebuilt only after barrier C of the next iteration (meanwhile the SM's other CTA builds its chunk)
            if (tid == 0 && !RDBG(2)) bulk_s2g(p.out + ((size_t)img * HW + q0) * Cout, st, (uint32_t)(npx * Cout * 4));
        } else {
            __syncthreads();
            float* const dst = p.out + ((size_t)img * HW + q0) * Cout;
            const int nf = npx * Cout;
            if (p.vec) {   // 16-byte aligned chunk: 128-bit streaming stores
                const float4* s4 = reinterpret_cast<const float4*>(st);
                float4* d4 = reinterpret_cast<float4*>(dst);
                for (int f = tid; f < (nf >> 2); f += kThreads) st_cs_f4(d4 + f, s4[f]);
            } else {
                for (int f = tid; f < nf; f += kThreads) dst[f] = st[f];
            }
            __syncthreads();   // the buffer is refilled two chunks later, but sobj / lists are reused by the next one
        }
        loaded_img = img;
        if (++ci == cpi) {
            ci = 0;
            ++img;
            ya = xa = 0;
        } else {
            ya += step_rows;
            xa += step_cols;
            if (xa >= W) {
                xa -= W;
                ++ya;
            }
        }
    }
    if (p.bulk && tid == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // stores complete before the CTA retires
}

int launch_render(const RenderParams& p0, int B, cudaStream_t st) {
    RenderParams p = p0;
    p.HW = p.H * p.W;
    int P = (kChunkBytes / (p.Cout * 4)) & ~3;
    // narrow outputs: 4096-pixel chunks keep the (object, row) items of a chunk short; the one-channel previous-frame
    // heatmap takes 64 KB chunks (0.121 -> 0.097 ms for the 512 images of BASELINE configs[3])
    const int p_cap = p.Cout == 1 ? 16384 : 4096;
    if (P > p_cap) P = p_cap;
    if (P > ((p.HW + 3) & ~3)) P = (p.HW + 3) & ~3;
    CVM_CHECK_ARG(P >= 4 && p.Cout <= kThreads, "render: %d output channels do not fit the staging buffer", p.Cout);
    p.P = P;
    p.cpi = (p.HW + P - 1) / P;
    p.n_chunks = (long long)B * p.cpi;
    if (p.n_chunks == 0) return CVM_OK;
    // bulk stores need 16-byte granules: base aligned, every image a whole number of them (chunks are P*Cout*4 bytes with
    // P % 4 == 0, the partial last chunk of an image then ends on a granule too)
    p.bulk = cvm_aligned16(p.out) && (((long long)p.HW * p.Cout) % 4 == 0);
    p.vec = p.bulk;
#ifdef CVM_EXPERIMENT
    if (const char* e = getenv("CVM_RENDER_SKIP")) p.dbg_skip = atoi(e);
    if RDBG(8) p.bulk = 0;   // experiment: plain stores instead of bulk copies
#endif
    const size_t smem = (size_t)kChunkBytes;
    CVM_SMEM_ATTR_ONCE((render_kernel), smem);
    long long grid = 2LL * cvm_num_sms();
    if (grid > p.n_chunks) grid = p.n_chunks;
    render_kernel<<<(unsigned)grid, kThreads, smem, st>>>(p);
    CVM_CHECK_LAUNCH("render_kernel");
    return CVM_OK;
}

}  // namespace

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
