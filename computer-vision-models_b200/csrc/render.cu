// Ground-truth render: gaussian splat (max-combine) + loss-weights (min-combine) + centre scatter + ignore areas.
//
// Replaces fill_heatmap (reference models/centernet/processor.py:17-38, a numba double loop per object) and the render
// part of ProcessImages.process (processor.py:264-334) for a whole batch; also gen_prev_heatmap
// (models/centertracker/processor.py:22-41) with n_planes = 1 and no weights plane.
//
// Bound: HBM writes.  Algorithmic bytes per image: 4*H*W*Ct, every byte written exactly once (zeros included).
//
// Structure: one CTA per tile of `rows` full image rows.  The (hm + 1) planes that can be non-trivial (heatmap
// channels + weights) live in shared memory as SoA planes; each thread OWNS a set of columns, so the max/min-combine
// over all objects of the image needs neither atomics nor barriers.  Gaussian argument and exp are evaluated in fp64
// (the reference does fp64 scalar math and stores fp32; B200 has full-rate-enough fp64 for ~1 exp per covered pixel).
// The tile is then composed to NHWC and streamed out with 128-bit stores; the handful of regression targets at centre
// pixels are patched afterwards by the same CTA.
#include "common.cuh"

namespace {

constexpr int kThreads = 256;   // power of two: column ownership is x & (kThreads - 1)
constexpr int kMaxObjSmem = 96;  // objects are processed in chunks of this many
constexpr int kMaxIgnSmem = 16;  // ignore boxes cached in shared memory per band (more are read from global)

struct RenderParams {
    const cvm_obj* objs;
    const int32_t* obj_offsets;
    const cvm_box* ignore;
    const int32_t* ign_offsets;
    float* out;        // [B,H,W,Cout]
    int H, W, Cout;
    int n_planes;      // heatmap planes kept in smem (hm)
    int wch;           // weights channel in the output pixel, -1 = none (prev-frame heatmap)
    int rows;          // rows per tile
    int tiles_per_band;  // tiles (of `rows` rows) handled by one CTA
    int bands_per_img;
    int plane_stride;  // rows*W + pad
    int per_class;     // 1: heat plane = obj.cls (Profile N); 0: plane 0 (reference as shipped)
    int force_explicit;
    int off_class, off_roff, off_box, off_track;
    int vec_ok;        // tile byte ranges are 16-byte aligned
    double R, alpha;
};

// derived per-object quantities, computed once per tile by one thread per object
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

__global__ void __launch_bounds__(kThreads, 3) render_kernel(const RenderParams p) {
    extern __shared__ __align__(16) float planes[];  // [n_planes + 1][plane_stride]; last = weights
    __shared__ ObjDerived sobj[kMaxObjSmem];
    __shared__ int s_list[kMaxObjSmem];
    __shared__ int s_nlist;

    const int tid = threadIdx.x;
    const int b = blockIdx.x / p.bands_per_img;
    const int band = blockIdx.x - b * p.bands_per_img;
    const int W = p.W, PS = p.plane_stride;
    const int Cout = p.Cout, hm = p.n_planes, wch = p.wch;
    float* const wplane = planes + (size_t)hm * PS;
    const int band_ya = band * p.tiles_per_band * p.rows;
    const int band_yb = min(p.H, band_ya + p.tiles_per_band * p.rows);

    const int o_begin = p.obj_offsets[b], o_end = p.obj_offsets[b + 1];
    float* const out_img = p.out + (size_t)b * p.H * W * Cout;
    const bool single_chunk = (o_end - o_begin) <= kMaxObjSmem;
    int loaded_base = -1;  // which chunk of objects currently sits in sobj
    if (single_chunk && o_end > o_begin) {  // the common case: derive once per band, reuse for every tile
        if (tid < o_end - o_begin) derive(p.objs[o_begin + tid], p, sobj[tid]);
        loaded_base = o_begin;
    }

    // compose mapping (vector path): the NHWC pattern repeats every 4 pixels = Cout float4s.  Thread t < NA handles float4
    // slot r = t % Cout of pixel groups g = t / Cout, + GS, ...; which plane/pixel feeds its 4 floats never changes, so the
    // loop is four shared loads through four pointers that advance by a fixed step (step 0 on a zero word for the
    // channels that are always zero) and one 128-bit store.
    __shared__ float s_zero[4];
    if (tid < 4) s_zero[tid] = 0.f;
    const int GS = kThreads / Cout, NA = GS * Cout;
    const int cg0 = tid / Cout;
    const float* csrc[4];
    int cinc[4];
    {
        const int cr = tid - cg0 * Cout;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int f = 4 * cr + k, pig = f / Cout, ch = f - pig * Cout;
            csrc[k] = s_zero;
            cinc[k] = 0;
            if (ch < hm || ch == wch) {
                csrc[k] = planes + (ch < hm ? ch : hm) * PS + pig + 4 * cg0;
                cinc[k] = 4 * GS;
            }
        }
    }
    // ignore boxes of this image, cached once per band (processor.py:318-323)
    __shared__ int s_ign[kMaxIgnSmem][4];   // sx, ex, sy, ey (clamped)
    int n_ign = 0, i_begin = 0;
    if (p.ignore != nullptr && wch >= 0) {
        i_begin = p.ign_offsets[b];
        n_ign = p.ign_offsets[b + 1] - i_begin;
        if (tid < min(n_ign, kMaxIgnSmem)) {
            const cvm_box bx = p.ignore[i_begin + tid];
            s_ign[tid][0] = max((int)bx.x, 0);
            s_ign[tid][1] = min(max((int)(bx.x + bx.w), 0), W);
            s_ign[tid][2] = max((int)bx.y, 0);
            s_ign[tid][3] = min(max((int)(bx.y + bx.h), 0), p.H);
        }
    }

    for (int ya = band_ya; ya < band_yb; ya += p.rows) {
        const int yb = min(band_yb, ya + p.rows);
        const int nrows = yb - ya;

        // ---- init: heat = 0, weights = 1 (processor.py:267-268) ----
        {
            float4* p4 = reinterpret_cast<float4*>(planes);
            const int n4 = (hm * PS) >> 2;
            for (int i = tid; i < n4; i += kThreads) p4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int i = (n4 << 2) + tid; i < hm * PS; i += kThreads) planes[i] = 0.f;
            for (int i = tid; i < PS; i += kThreads) wplane[i] = 1.f;
            if (tid == 0) s_nlist = 0;
        }
        __syncthreads();  // planes are handed over from linear-index owners to column owners

        // ---- splat: each thread owns columns x with x % 256 == tid for all rows of the tile ----
        for (int base = o_begin; base < o_end; base += kMaxObjSmem) {
            const int n = min(kMaxObjSmem, o_end - base);
            if (loaded_base != base) {  // only with more than one chunk of objects
                __syncthreads();
                if (tid < n) derive(p.objs[base + tid], p, sobj[tid]);
                loaded_base = base;
                if (tid == 0) s_nlist = 0;
                __syncthreads();
            }
            if (tid < n) {  // objects whose window touches this tile (order is irrelevant for max/min)
                const ObjDerived& d = sobj[tid];
                if (max(d.y0, ya) < min(d.y1, yb) && d.x0 < d.x1) s_list[atomicAdd(&s_nlist, 1)] = tid;
            }
            __syncthreads();
            const int nl = s_nlist;
            for (int li = 0; li < nl; ++li) {
                const ObjDerived& d = sobj[s_list[li]];
                const int r0 = max(d.y0, ya), r1 = min(d.y1, yb);
                float* const hp = planes + (size_t)d.plane * PS;
                for (int x = d.x0 + ((tid - d.x0) & (kThreads - 1)); x < d.x1; x += kThreads) {
                    const double dx = (double)(x - d.cx);
                    const double ax = dx * dx * d.inv2vx;
                    for (int y = r0; y < r1; ++y) {
                        const double dy = (double)(y - d.cy);
                        const double g = exp(-(ax + dy * dy * d.inv2vy));                 // processor.py:34-36
                        const int idx = (y - ya) * W + x;
                        hp[idx] = fmaxf(hp[idx], (float)(g * d.peak));                    // :37
                        if (wch >= 0) wplane[idx] = fminf(wplane[idx], (float)(1.0 - d.rw * g));   // :38
                    }
                }
            }
        }

        // ---- ignore areas: weights = 0, input-px numbers used as mask indices (processor.py:318-323) ----
        for (int i = 0; i < n_ign; ++i) {
            int sx, ex, sy, ey;
            if (i < kMaxIgnSmem) {
                sx = s_ign[i][0], ex = s_ign[i][1], sy = s_ign[i][2], ey = s_ign[i][3];
            } else {
                const cvm_box bx = p.ignore[i_begin + i];
                sx = max((int)bx.x, 0), ex = min(max((int)(bx.x + bx.w), 0), W);
                sy = max((int)bx.y, 0), ey = min(max((int)(bx.y + bx.h), 0), p.H);
            }
            const int r0 = max(sy, ya), r1 = min(ey, yb);
            for (int x = sx + ((tid - sx) & (kThreads - 1)); x < ex; x += kThreads)   // same column ownership as the splat
                for (int y = r0; y < r1; ++y) wplane[(y - ya) * W + x] = 0.f;
        }
        __syncthreads();

        // ---- compose NHWC and stream out ----
        const int tile_floats = nrows * W * Cout;
        float* const out_tile = out_img + (size_t)ya * W * Cout;
        if (p.vec_ok && ((nrows * W) & 3) == 0) {  // a ragged last tile may not be whole groups of 4 pixels
            if (tid < NA) {
                const int n_groups = (nrows * W) >> 2;
                float4* dst = reinterpret_cast<float4*>(out_tile) + tid;
                const float *s0 = csrc[0], *s1 = csrc[1], *s2 = csrc[2], *s3 = csrc[3];
                for (int g = cg0; g < n_groups; g += GS, dst += NA) {
                    st_cs_f4(dst, make_float4(*s0, *s1, *s2, *s3));
                    s0 += cinc[0];
                    s1 += cinc[1];
                    s2 += cinc[2];
                    s3 += cinc[3];
                }
            }
        } else {
            for (int f = tid; f < tile_floats; f += kThreads) {
                const int px = f / Cout, ch = f - px * Cout;
                out_tile[f] = ch < hm ? planes[(size_t)ch * PS + px] : (ch == wch ? wplane[px] : 0.f);
            }
        }
        __syncthreads();  // the planes are re-initialised for the next tile
    }

    // ---- centre scatter (processor.py:288-299): patched after the band is out; last object wins per pixel ----
    if (p.off_class < 0 && p.off_roff < 0 && p.off_box < 0 && p.off_track < 0) return;
    for (int base = o_begin; base < o_end; base += kMaxObjSmem) {
        const int n = min(kMaxObjSmem, o_end - base);
        if (loaded_base != base) {
            __syncthreads();
            if (tid < n) derive(p.objs[base + tid], p, sobj[tid]);
            loaded_base = base;
            __syncthreads();
        }
        if (tid < n) {
            const ObjDerived& d = sobj[tid];
            if (d.scy >= band_ya && d.scy < band_yb && d.scx >= 0) {
                // a later object (list order) at the same pixel overwrites r_offset/fullbox/track
                bool last = true;
                for (int j = base + tid + 1; j < o_end && last; ++j) {
                    if (j - base < n) {
                        const ObjDerived& e = sobj[j - base];
                        if (e.scx == d.scx && e.scy == d.scy) last = false;
                    } else {
                        ObjDerived e;
                        derive(p.objs[j], p, e);
                        if (e.scx == d.scx && e.scy == d.scy) last = false;
                    }
                }
                float* px = out_img + ((size_t)d.scy * W + d.scx) * Cout;
                if (p.off_class >= 0 && d.cls >= 0 && p.off_class + d.cls < Cout) px[p.off_class + d.cls] = 1.0f;   // :290
                if (last) {
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
                }
            }
        }
    }
}

int pick_rows(int H, int W, int n_planes_total, size_t* smem_bytes, int* plane_stride) {
    // as many rows as keep the planes <= ~36 KB (>= 4 CTAs/SM by shared memory), at least one
    int rows = 1;
    const size_t budget = 36 * 1024;
    while (rows < H && (size_t)(rows * 2) * W * n_planes_total * 4 <= budget) rows *= 2;
    *plane_stride = ((rows * W + 3) & ~3) + 1;   // odd stride: plane-to-plane bank offset of 1
    *smem_bytes = (size_t)(*plane_stride) * n_planes_total * 4 + 16;
    return rows;
}

int launch_render(const RenderParams& p0, int B, cudaStream_t st) {
    RenderParams p = p0;
    size_t smem = 0;
    p.rows = pick_rows(p.H, p.W, p.n_planes + 1, &smem, &p.plane_stride);
    if (smem > 200 * 1024) {
        cvm_set_error("render: one row of %d planes x %d px does not fit in shared memory", p.n_planes + 1, p.W);
        return CVM_ERR_ARG;
    }
    // a CTA renders a band of consecutive tiles of one image (object records are derived once per band); bands are sized so
    // that the grid still fills the machine ~4x over
    const int tiles_per_img = (p.H + p.rows - 1) / p.rows;
    int tpb = 8;
    const long long want = 4LL * cvm_num_sms() * 4;
    while (tpb > 1 && (long long)B * ((tiles_per_img + tpb - 1) / tpb) < want) tpb >>= 1;
    p.tiles_per_band = tpb;
    p.bands_per_img = (tiles_per_img + tpb - 1) / tpb;
    // 128-bit stores need every tile start (and the image start) 16-byte aligned and whole groups of 4 pixels per tile
    p.vec_ok = cvm_aligned16(p.out) && (((long long)p.rows * p.W) % 4 == 0) && (((long long)p.H * p.W * p.Cout) % 4 == 0) &&
               p.Cout <= kThreads;
    CVM_CHECK_CUDA(cudaFuncSetAttribute(render_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const long long grid = (long long)B * p.bands_per_img;
    if (grid == 0) return CVM_OK;
    CVM_CHECK_ARG(grid < 2147483647LL, "render grid too large");
    render_kernel<<<(unsigned)grid, kThreads, smem, st>>>(p);
    CVM_CHECK_LAUNCH("render_kernel");
    return CVM_OK;
}

}  // namespace

extern "C" int cvm_render_gt(const cvm_layout* L, const cvm_obj* objs, const int32_t* obj_offsets, const cvm_box* ignore,
                             const int32_t* ign_offsets, int B, float* y_true, void* stream) {
    CVM_CHECK_ARG(L && obj_offsets && y_true, "NULL pointer argument");
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
