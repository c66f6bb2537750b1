// Batched pre-processing front-end: raw labelled boxes -> the object records and ignore boxes the render consumes,
// entirely on the device, so that y_true is born there without the per-sample host loop of the reference
// (data/base_data_generator.py:29-48 calling ProcessImages.process once per sample).
//
// Per object, exactly the reference's fp64 arithmetic (models/centernet/processor.py):
//   clip_to_img (:46-56): x' = clip(x, 0, W), mx' = clip(x + w, 0, W), w' = mx' - x' (same in y)
//   filter (:241-253):    w' * h' > MIN_BOX_AREA -> object (list order kept), else -> ignore area
// The work is tiny (a few thousand objects): one CTA, one thread per image, an order-preserving compaction with a
// shared-memory prefix sum over the images of a tile.
#include "common.cuh"

namespace {

constexpr int kThreads = 256;

struct PrepParams {
    const double* boxes;         // [n,4] x, y, w, h in input px
    const int32_t* cls;          // [n] class index (OD_CLASS_IDX[obj_class])
    const float* track;          // [n,2] track_offset targets or NULL
    const int32_t* raw_offsets;  // [B+1]
    int B;
    double img_w, img_h, min_area;
    cvm_obj* objs;
    int32_t* obj_offsets;        // [B+1]
    cvm_box* ignore;
    int32_t* ign_offsets;        // [B+1]
};

__device__ __forceinline__ double clipd(double v, double lo, double hi) { return fmin(fmax(v, lo), hi); }

__global__ void __launch_bounds__(kThreads) prepare_objects_kernel(const PrepParams p) {
    __shared__ int s_keep[kThreads], s_ign[kThreads];
    __shared__ int s_base[2];
    const int tid = threadIdx.x;
    if (tid == 0) {
        s_base[0] = s_base[1] = 0;
        p.obj_offsets[0] = 0;
        p.ign_offsets[0] = 0;
    }
    __syncthreads();
    for (int b0 = 0; b0 < p.B; b0 += kThreads) {
        const int b = b0 + tid;
        int n_keep = 0, n_ign = 0, r0 = 0, r1 = 0;
        if (b < p.B) {
            r0 = p.raw_offsets[b];
            r1 = p.raw_offsets[b + 1];
            for (int i = r0; i < r1; ++i) {
                const double x = p.boxes[4 * i], y = p.boxes[4 * i + 1], w = p.boxes[4 * i + 2], h = p.boxes[4 * i + 3];
                const double x0 = clipd(x, 0.0, p.img_w), y0 = clipd(y, 0.0, p.img_h);
                const double cw = clipd(x + w, 0.0, p.img_w) - x0, ch = clipd(y + h, 0.0, p.img_h) - y0;
                if (cw * ch > p.min_area) ++n_keep;
                else ++n_ign;
            }
        }
        s_keep[tid] = n_keep;
        s_ign[tid] = n_ign;
        __syncthreads();
        // exclusive prefix sums over the tile (Hillis-Steele on two arrays)
        for (int o = 1; o < kThreads; o <<= 1) {
            const int a = tid >= o ? s_keep[tid - o] : 0, c = tid >= o ? s_ign[tid - o] : 0;
            __syncthreads();
            s_keep[tid] += a;
            s_ign[tid] += c;
            __syncthreads();
        }
        const int base_k = s_base[0], base_i = s_base[1];
        int ok = base_k + s_keep[tid] - n_keep, oi = base_i + s_ign[tid] - n_ign;
        if (b < p.B) {
            p.obj_offsets[b + 1] = base_k + s_keep[tid];
            p.ign_offsets[b + 1] = base_i + s_ign[tid];
            for (int i = r0; i < r1; ++i) {
                const double x = p.boxes[4 * i], y = p.boxes[4 * i + 1], w = p.boxes[4 * i + 2], h = p.boxes[4 * i + 3];
                const double x0 = clipd(x, 0.0, p.img_w), y0 = clipd(y, 0.0, p.img_h);
                const double cw = clipd(x + w, 0.0, p.img_w) - x0, ch = clipd(y + h, 0.0, p.img_h) - y0;
                if (cw * ch > p.min_area) {
                    cvm_obj o;
                    o.x = x0;
                    o.y = y0;
                    o.w = cw;
                    o.h = ch;
                    o.cx = o.cy = 0;
                    o.cls = p.cls[i];
                    o.flags = 0;
                    o.peak = 1.0f;
                    o.track[0] = p.track ? p.track[2 * i] : 0.f;
                    o.track[1] = p.track ? p.track[2 * i + 1] : 0.f;
                    o._pad = 0.f;
                    p.objs[ok++] = o;
                } else {
                    cvm_box g;
                    g.x = x0;
                    g.y = y0;
                    g.w = cw;
                    g.h = ch;
                    p.ignore[oi++] = g;
                }
            }
        }
        __syncthreads();
        if (tid == kThreads - 1) {
            s_base[0] = base_k + s_keep[tid];
            s_base[1] = base_i + s_ign[tid];
        }
        __syncthreads();
    }
}

}  // namespace

extern "C" int cvm_prepare_objects(const double* raw_boxes, const int32_t* raw_cls, const float* raw_track,
                                   const int32_t* raw_offsets, int B, double img_w, double img_h, double min_box_area,
                                   cvm_obj* objs, int32_t* obj_offsets, cvm_box* ignore, int32_t* ign_offsets, void* stream) {
    CVM_CHECK_ARG(raw_offsets && obj_offsets && ign_offsets && B >= 0, "bad argument");
    // raw_boxes / raw_cls may be NULL when the batch holds no object at all (the offsets say so on the device)
    CVM_CHECK_ARG(B == 0 || (objs && ignore), "NULL pointer argument");
    PrepParams p;
    p.boxes = raw_boxes;
    p.cls = raw_cls;
    p.track = raw_track;
    p.raw_offsets = raw_offsets;
    p.B = B;
    p.img_w = img_w;
    p.img_h = img_h;
    p.min_area = min_box_area;
    p.objs = objs;
    p.obj_offsets = obj_offsets;
    p.ignore = ignore;
    p.ign_offsets = ign_offsets;
    prepare_objects_kernel<<<1, kThreads, 0, reinterpret_cast<cudaStream_t>(stream)>>>(p);
    CVM_CHECK_LAUNCH("prepare_objects_kernel");
    return CVM_OK;
}
