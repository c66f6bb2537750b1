// Profile-R decode (the reference as shipped) and the semseg colour/argmax step.
//
// cvm_decode_window9 replaces process_2d_output (reference models/centernet/post_processing.py:6-66): a pure-Python
// loop of (H-8)(W-8) np.argmax calls per image.  Three small kernels: (1) pull the objectness channel out of the NHWC
// tensor into a compact plane (this is the one pass over y_pred: HBM-bound, 4*H*W*stride bytes per image), (2) window
// test on the L2-resident plane with an early-out on the confidence threshold, (3) one CTA per image compacts the hits
// in scan order (the reference's output order) and assembles class / centre / box.
//
// cvm_semseg_argmax replaces to_3channel (reference common/utils/image.py:72-100).
#include <limits.h>

#include "common.cuh"

namespace {

__global__ void __launch_bounds__(256) extract_plane_kernel(const float* __restrict__ yp, int stride, long long n_pixels,
                                                            float* __restrict__ plane) {
    const long long step = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_pixels; i += step)
        plane[i] = __ldg(yp + i * stride);
}

// centre pixel is the FIRST argmax (row-major) of its window and strictly above the threshold
// (post_processing.py:30-35): strictly greater than every earlier element, >= every later one.
__global__ void __launch_bounds__(256) window_peak_kernel(const float* __restrict__ plane, int B, int H, int W, int r,
                                                          float min_conf, unsigned char* __restrict__ mask) {
    const long long n = (long long)B * H * W;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int x = (int)(i % W);
    const int y = (int)((i / W) % H);
    unsigned char hit = 0;
    if (y >= r && y < H - r && x >= r && x < W - r) {
        const float v = plane[i];
        if (v > min_conf) {
            hit = 1;
            for (int dy = -r; dy <= r && hit; ++dy) {
                const float* row = plane + i + (long long)dy * W;
                for (int dx = -r; dx <= r; ++dx) {
                    const float o = row[dx];
                    const bool before = (dy < 0) || (dy == 0 && dx < 0);
                    if (before ? !(o < v) : (o > v)) {
                        hit = 0;
                        break;
                    }
                }
            }
        }
    }
    mask[i] = hit;
}

struct EmitParams {
    const float* yp;
    const float* plane;
    const unsigned char* mask;
    int stride, H, W, nb_classes;
    int off_class, off_roff, off_box;
    float R;
    const cvm_roi* rois;
    int max_out;
    int32_t* counts;
    int32_t* cls;
    int32_t* pix;
    float* scores;
    float* centers;
    float* boxes;
};

__global__ void __launch_bounds__(256) window_emit_kernel(const EmitParams p) {
    __shared__ int s_cnt[256];
    const int b = blockIdx.x, tid = threadIdx.x;
    const int HW = p.H * p.W;
    const int chunk = (HW + 255) / 256;
    const int beg = min(HW, tid * chunk), end = min(HW, beg + chunk);
    const unsigned char* m = p.mask + (size_t)b * HW;
    int c = 0;
    for (int i = beg; i < end; ++i) c += m[i];
    s_cnt[tid] = c;
    __syncthreads();
    // exclusive scan over 256 chunk counts (Hillis-Steele in shared memory)
    for (int o = 1; o < 256; o <<= 1) {
        const int v = tid >= o ? s_cnt[tid - o] : 0;
        __syncthreads();
        s_cnt[tid] += v;
        __syncthreads();
    }
    int pos = s_cnt[tid] - c;
    if (tid == 255) p.counts[b] = s_cnt[255];
    const cvm_roi roi = p.rois ? p.rois[b] : cvm_roi{1.0f, 0.0f, 0.0f, 0.0f};
    for (int i = beg; i < end; ++i) {
        if (!m[i]) continue;
        if (pos < p.max_out) {
            const int y = i / p.W, x = i - y * p.W;
            const float* px = p.yp + ((size_t)b * HW + i) * p.stride;
            int ci = 0;
            if (p.off_class >= 0) {  // np.argmax: first max (post_processing.py:39-41)
                float best = px[p.off_class];
                for (int k = 1; k < p.nb_classes; ++k) {
                    const float v = px[p.off_class + k];
                    if (v > best) {
                        best = v;
                        ci = k;
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
            const float cx = __fsub_rn(__fmul_rn(roi.inv_scale, __fmul_rn(__fadd_rn((float)x, dx), p.R)), roi.off_left);
            const float cy = __fsub_rn(__fmul_rn(roi.inv_scale, __fmul_rn(__fadd_rn((float)y, dy), p.R)), roi.off_top);
            const float bw = __fmul_rn(w, roi.inv_scale), bh = __fmul_rn(h, roi.inv_scale);
            const size_t o = (size_t)b * p.max_out + pos;
            p.cls[o] = ci;
            p.pix[o] = i;
            p.scores[o] = p.plane[(size_t)b * HW + i];
            p.centers[o * 2 + 0] = cx;
            p.centers[o * 2 + 1] = cy;
            p.boxes[o * 4 + 0] = __fsub_rn(cx, __fmul_rn(bw, 0.5f));
            p.boxes[o * 4 + 1] = __fsub_rn(cy, __fmul_rn(bh, 0.5f));
            p.boxes[o * 4 + 2] = bw;
            p.boxes[o * 4 + 3] = bh;
        }
        ++pos;
    }
}

constexpr int kMaxCls = 32;

struct ArgmaxParams {
    const float* in;
    long long n_pixels;
    int stride, off, n_cls, mode, apply_softmax, use_weight, has_thr;
    double thr;
    const unsigned char* lut;
    unsigned char* out;
};

__global__ void __launch_bounds__(256) semseg_argmax_kernel(const ArgmaxParams p) {
    const long long step = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < p.n_pixels; i += step) {
        const float* px = p.in + i * p.stride + p.off;
        float a[kMaxCls];
        float mn = __ldg(px);
        for (int k = 0; k < p.n_cls; ++k) {
            a[k] = __ldg(px + k);
            mn = fminf(mn, a[k]);
        }
        if (p.mode == 0) {  // plain class ids
            int idx = 0;
            for (int k = 1; k < p.n_cls; ++k)
                if (a[k] > a[idx]) idx = k;
            p.out[i] = (unsigned char)idx;
            continue;
        }
        if (p.apply_softmax) {  // image.py:81-83: shift by the min, divide by the (sequential fp32) sum
            float s = 0.f;
            for (int k = 0; k < p.n_cls; ++k) {
                a[k] = __fsub_rn(a[k], mn);
                s = __fadd_rn(s, a[k]);
            }
            for (int k = 0; k < p.n_cls; ++k) a[k] = __fdiv_rn(a[k], s);
        }
        int idx = 0;
        float score;
        if (p.n_cls == 1) {
            score = a[0];  // image.py:85-87
        } else {
            // np.argmax: first maximum; a NaN counts as the maximum (first NaN wins)
            bool nan_found = isnan(a[0]);
            for (int k = 1; k < p.n_cls && !nan_found; ++k) {
                if (isnan(a[k])) {
                    idx = k;
                    nan_found = true;
                } else if (a[k] > a[idx]) {
                    idx = k;
                }
            }
            const float v = a[idx];
            score = isnan(v) ? 0.f : fminf(1.f, fmaxf(0.f, v));  // numba: min(1.0, max(0.0, nan)) == 0.0
        }
        unsigned char c0 = 0, c1 = 0, c2 = 0;
        const double sd = (double)score;
        if (!p.has_thr || sd > p.thr) {  // image.py:92 (NaN fails the compare)
            const double wgt = p.use_weight ? sd : 1.0;
            const double v0 = wgt * (double)p.lut[idx * 3 + 0], v1 = wgt * (double)p.lut[idx * 3 + 1],
                         v2 = wgt * (double)p.lut[idx * 3 + 2];
            c0 = isnan(v0) ? 0 : (unsigned char)(int)v0;  // truncating cast (image.py:99)
            c1 = isnan(v1) ? 0 : (unsigned char)(int)v1;
            c2 = isnan(v2) ? 0 : (unsigned char)(int)v2;
        }
        p.out[i * 3 + 0] = c0;
        p.out[i * 3 + 1] = c1;
        p.out[i * 3 + 2] = c2;
    }
}

}  // namespace

extern "C" size_t cvm_decode_window9_workspace_bytes(const cvm_layout* L, int B) {
    if (!L || B < 0) return 0;
    const size_t n = (size_t)B * L->H * L->W;
    return ((n * 4 + 255) & ~(size_t)255) + n + 256;
}

extern "C" int cvm_decode_window9(const cvm_layout* L, const float* y_pred, int pred_stride, int B, int window,
                                  float min_conf, const cvm_roi* rois, int max_out, int32_t* counts, int32_t* cls,
                                  int32_t* pix, float* scores, float* centers, float* boxes, void* ws, size_t ws_bytes,
                                  void* stream) {
    CVM_CHECK_ARG(L && y_pred && counts && cls && pix && scores && centers && boxes && ws, "NULL pointer argument");
    CVM_CHECK_ARG(B >= 0 && L->H > 0 && L->W > 0 && pred_stride >= L->Cp, "bad shape");
    CVM_CHECK_ARG(window >= 1 && (window & 1) && window <= 31, "window must be odd and <= 31");
    CVM_CHECK_ARG(max_out >= 1, "max_out < 1");
    CVM_CHECK_ARG((long long)L->H * L->W < 2147483647LL, "image too large");
    if (ws_bytes < cvm_decode_window9_workspace_bytes(L, B)) {
        cvm_set_error("workspace too small");
        return CVM_ERR_WS;
    }
    if (B == 0) return CVM_OK;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const long long n = (long long)B * L->H * L->W;
    float* plane = static_cast<float*>(ws);
    unsigned char* mask = static_cast<unsigned char*>(ws) + (((size_t)n * 4 + 255) & ~(size_t)255);
    long long g = (n + 255) / 256;
    const long long gmax = (long long)cvm_num_sms() * 16;
    extract_plane_kernel<<<(unsigned)(g < gmax ? g : gmax), 256, 0, st>>>(y_pred, pred_stride, n, plane);
    CVM_CHECK_LAUNCH("extract_plane_kernel");
    CVM_CHECK_ARG(g < 2147483647LL, "grid too large");
    window_peak_kernel<<<(unsigned)g, 256, 0, st>>>(plane, B, L->H, L->W, window / 2, min_conf, mask);
    CVM_CHECK_LAUNCH("window_peak_kernel");
    EmitParams e;
    memset(&e, 0, sizeof(e));
    e.yp = y_pred;
    e.plane = plane;
    e.mask = mask;
    e.stride = pred_stride;
    e.H = L->H;
    e.W = L->W;
    e.nb_classes = L->nb_classes;
    e.off_class = L->off_class;
    e.off_roff = L->off_roff;
    e.off_box = L->off_box;
    e.R = (float)L->R;
    e.rois = rois;
    e.max_out = max_out;
    e.counts = counts;
    e.cls = cls;
    e.pix = pix;
    e.scores = scores;
    e.centers = centers;
    e.boxes = boxes;
    window_emit_kernel<<<B, 256, 0, st>>>(e);
    CVM_CHECK_LAUNCH("window_emit_kernel");
    return CVM_OK;
}

extern "C" int cvm_semseg_argmax(const float* in, long long n_pixels, int stride, int off, int n_cls, int mode,
                                 int apply_softmax, int use_weight, double threshold, const unsigned char* lut_bgr,
                                 unsigned char* out, void* stream) {
    CVM_CHECK_ARG(in && out, "NULL pointer argument");
    CVM_CHECK_ARG(n_pixels >= 0 && stride >= 1 && off >= 0 && n_cls >= 1 && off + n_cls <= stride, "bad shape");
    CVM_CHECK_ARG(n_cls <= kMaxCls, "n_cls=%d above %d", n_cls, kMaxCls);
    CVM_CHECK_ARG(mode == 0 || (mode == 1 && lut_bgr), "mode must be 0 (ids) or 1 (BGR, needs lut)");
    if (n_pixels == 0) return CVM_OK;
    ArgmaxParams p;
    memset(&p, 0, sizeof(p));
    p.in = in;
    p.n_pixels = n_pixels;
    p.stride = stride;
    p.off = off;
    p.n_cls = n_cls;
    p.mode = mode;
    p.apply_softmax = apply_softmax;
    p.use_weight = use_weight;
    p.has_thr = !(threshold != threshold);
    p.thr = threshold;
    p.lut = lut_bgr;
    p.out = out;
    long long g = (n_pixels + 255) / 256;
    const long long gmax = (long long)cvm_num_sms() * 16;
    semseg_argmax_kernel<<<(unsigned)(g < gmax ? g : gmax), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(p);
    CVM_CHECK_LAUNCH("semseg_argmax_kernel");
    return CVM_OK;
}

// ---- CenterTrack association -------------------------------------------------------------------------
namespace {

struct AssocParams {
    const float* centers;
    const float* track;
    const float* boxes;
    const float* scores;
    const int32_t* cls;
    const float* prev_centers;
    const float* prev_sizes;
    const int32_t* prev_cls;
    const int32_t* prev_count;
    int32_t* match;
    int K, M;
    float min_score;
};

// One warp per image.  The greedy loop over the detections is sequential by definition (a match removes a previous track
// for all later detections); the search over the previous tracks is spread over the lanes and reduced with shuffles.
// Un-fused fp32 ops, x term + y term: the distances are the ones numpy computes, so ties break identically.
__global__ void __launch_bounds__(32) track_associate_kernel(const AssocParams p) {
    extern __shared__ __align__(16) unsigned char assoc_smem[];
    const int b = blockIdx.x, lane = threadIdx.x, K = p.K, M = p.M;
    float* pcx = reinterpret_cast<float*>(assoc_smem);
    float* pcy = pcx + M;
    float* psz = pcy + M;
    int* pcl = reinterpret_cast<int*>(psz + M);   // class of the previous track, INT_MIN once it is taken
    int n_prev = p.prev_count ? p.prev_count[b] : M;
    n_prev = n_prev < 0 ? 0 : (n_prev > M ? M : n_prev);
    for (int j = lane; j < n_prev; j += 32) {
        const size_t o = (size_t)b * M + j;
        pcx[j] = p.prev_centers[o * 2];
        pcy[j] = p.prev_centers[o * 2 + 1];
        psz[j] = __fmul_rn(p.prev_sizes[o * 2], p.prev_sizes[o * 2 + 1]);
        pcl[j] = p.prev_cls[o];
    }
    __syncwarp();
    for (int i = 0; i < K; ++i) {
        const size_t o = (size_t)b * K + i;
        int found = -1;
        const float s = p.scores[o];
        if (s >= p.min_score && p.cls[o] != INT_MIN) {   // (false for NaN; INT_MIN marks taken tracks) - uniform over the lanes
            const float qx = __fadd_rn(p.centers[o * 2], p.track[o * 2]), qy = __fadd_rn(p.centers[o * 2 + 1], p.track[o * 2 + 1]);
            const float item = __fmul_rn(p.boxes[o * 4 + 2], p.boxes[o * 4 + 3]);
            const int c = p.cls[o];
            float best = __int_as_float(0x7f800000);
            int bj = INT_MAX;
            for (int j = lane; j < n_prev; j += 32) {
                if (pcl[j] != c) continue;
                const float dx = __fsub_rn(pcx[j], qx), dy = __fsub_rn(pcy[j], qy);
                const float d = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
                if (d > psz[j] || d > item || !(d < best)) continue;   // (strict <: a lane keeps its first minimum)
                best = d;
                bj = j;
            }
#pragma unroll
            for (int sft = 16; sft > 0; sft >>= 1) {
                const float ob = __shfl_xor_sync(0xffffffffu, best, sft);
                const int oj = __shfl_xor_sync(0xffffffffu, bj, sft);
                if (oj != INT_MAX && (bj == INT_MAX || ob < best || (ob == best && oj < bj))) {
                    best = ob;
                    bj = oj;
                }
            }
            if (bj != INT_MAX) {
                found = bj;
                if (lane == 0) pcl[bj] = INT_MIN;   // taken
            }
            __syncwarp();
        }
        if (lane == 0) p.match[o] = found;
    }
}

}  // namespace

extern "C" int cvm_track_associate(const float* centers, const float* track, const float* boxes, const float* scores, const int32_t* cls,
                                   int B, int K, float min_score, const float* prev_centers, const float* prev_sizes,
                                   const int32_t* prev_cls, const int32_t* prev_count, int M, int32_t* match, void* stream) {
    CVM_CHECK_ARG(B >= 0 && K >= 0 && M >= 0, "bad shape B=%d K=%d M=%d", B, K, M);
    if (B == 0 || K == 0) return CVM_OK;
    CVM_CHECK_ARG(centers && track && boxes && scores && cls && match, "NULL pointer argument");
    CVM_CHECK_ARG(M == 0 || (prev_centers && prev_sizes && prev_cls), "NULL previous-frame argument");
    const size_t smem = (size_t)M * 16;
    CVM_CHECK_ARG(smem <= 200 * 1024, "M=%d previous tracks per image do not fit shared memory", M);
    AssocParams p;
    p.centers = centers;
    p.track = track;
    p.boxes = boxes;
    p.scores = scores;
    p.cls = cls;
    p.prev_centers = prev_centers;
    p.prev_sizes = prev_sizes;
    p.prev_cls = prev_cls;
    p.prev_count = prev_count;
    p.match = match;
    p.K = K;
    p.M = M;
    p.min_score = min_score;
    if (smem > 48 * 1024) CVM_SMEM_ATTR_ONCE((track_associate_kernel), smem);
    track_associate_kernel<<<B, 32, smem, reinterpret_cast<cudaStream_t>(stream)>>>(p);
    CVM_CHECK_LAUNCH("track_associate_kernel");
    return CVM_OK;
}
