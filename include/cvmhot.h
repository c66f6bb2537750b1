/*
 * cvmhot.h — C ABI of libcvmhot.so: the B200 (sm_100a) CenterNet/CenterTracker heatmap hot path.
 *
 * The reference (j-o-d-o/computer-vision-models) is pure Python and has no FFI of its own; the "operator API"
 * of this path is a handful of Python call sites.  Each entry point below replaces the tensor math behind one
 * of them (paths relative to the reference root):
 *
 *   cvm_prepare_objects  clip_to_img + MIN_BOX_AREA filter of ProcessImages.process (processor.py:46-56,241-253), batched
 *   cvm_render_gt        fill_heatmap (models/centernet/processor.py:17-38) + the render part of
 *                        ProcessImages.process (processor.py:264-334) incl. calc_img_data's centre math (:58-67)
 *   cvm_render_prev_hm   CenterTrackerProcess.gen_prev_heatmap (models/centertracker/processor.py:22-41)
 *   cvm_loss_fwd         CenternetLoss.obj_focal_loss / calc_loss / call (models/centernet/loss.py:31-60,107-155),
 *                        CentertrackerLoss.track_offset_loss (models/centertracker/loss.py:16-28),
 *                        MultitaskLoss.calc_centernet's channel slicing (models/multitask/loss.py:21-24,44-47)
 *   cvm_loss_finalize    the tf.cond normalisation + weighting of the same functions (loss.py:59,130,140-153)
 *   cvm_loss_bwd         d(total)/d(y_pred) of the above (TF autograd in the reference)
 *   cvm_decode_topk      north_star's canonical CenterNet decode (3x3 NMS, per-image top-K, head gather, box assembly
 *                        with the fullbox math of models/centernet/post_processing.py:43-52)
 *   cvm_decode_window9   process_2d_output as shipped (post_processing.py:6-66): 9x9 first-argmax + threshold
 *   cvm_semseg_argmax    to_3channel (common/utils/image.py:72-100)
 *   cvm_track_associate   greedy CenterTrack association of cvm_decode_topk's centre + track_offset with
 *                        the previous frame's centres; the reference stops at the loss (models/centertracker/__init__.py:5
 *                        exports params / loss / processor only), so this follows the published algorithm
 *
 * Conventions: all tensors are NHWC, fp32, C-contiguous DEVICE pointers owned by the caller; nothing is allocated,
 * freed or synchronised inside; every call only enqueues work on `stream` (a cudaStream_t / CUstream passed as void*).
 * Scratch memory comes from the caller (`ws`, at least the matching *_workspace_bytes()).  Calls are thread-safe for
 * concurrent use on different streams with different workspaces.  Return value: 0 = OK, negative = error code;
 * cvm_last_error() returns a thread-local message.
 */
#ifndef CVMHOT_H
#define CVMHOT_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CVM_MAX_FIELDS 8
#define CVM_NPART 16              /* length of the loss partials vector (fp64) */

/* loss kinds of CenternetLoss.calc_loss (loss.py:115-124) */
#define CVM_KIND_MSE 0
#define CVM_KIND_MAE 1
#define CVM_KIND_MAPE 2
#define CVM_KIND_CE 3
/* post transform applied in finalize (orientation_loss, loss.py:98) */
#define CVM_POST_NONE 0
#define CVM_POST_ORIENT 1

#define CVM_OK 0
#define CVM_ERR_ARG (-1)          /* bad argument / unsupported layout */
#define CVM_ERR_ALIGN (-2)        /* pointer not 16-byte aligned */
#define CVM_ERR_WS (-3)           /* workspace too small */
#define CVM_ERR_CUDA (-4)         /* a CUDA call failed (message has the CUDA error string) */

/* Channel layout of y_pred [.., Cp] / y_true [.., Ct = Cp + 1] (params.py:45-77; weights plane last, processor.py:334). */
typedef struct cvm_layout {
    int32_t H, W;                 /* mask size */
    int32_t hm;                   /* leading heatmap channels: 1 (reference as shipped) or nb_classes (canonical) */
    int32_t nb_classes;
    int32_t Cp, Ct;
    int32_t off_class;            /* start of the class-logit field, -1 if inactive */
    int32_t off_roff;             /* r_offset (x, y), -1 if inactive */
    int32_t off_box;              /* fullbox (w, h in input px), -1 if inactive */
    int32_t off_track;            /* track_offset (x, y), -1 if inactive */
    int32_t n_fields;             /* loss terms, in the order of loss.py:142-153 */
    int32_t field_off[CVM_MAX_FIELDS];
    int32_t field_size[CVM_MAX_FIELDS];
    int32_t field_kind[CVM_MAX_FIELDS];
    int32_t field_post[CVM_MAX_FIELDS];
    float field_weight[CVM_MAX_FIELDS];
    float focal_a, focal_b;       /* FOCAL_LOSS_ALPHA / BETA (params.py:40-41) */
    double R;                     /* params.R (params.py:34) */
    double alpha;                 /* VARIANCE_ALPHA (params.py:35) */
} cvm_layout;

#define CVM_OBJ_EXPLICIT_CENTER 1 /* use cx,cy as given (fill_heatmap drop-in / prev-frame heatmap) */
#define CVM_OBJ_NO_SCATTER 2      /* do not write the regression targets at the centre pixel */

/* One object; 64 bytes.  x,y,w,h is the clipped bbox in INPUT px exactly as the reference holds it (fp64). */
typedef struct cvm_obj {
    double x, y, w, h;
    int32_t cx, cy;               /* only read with CVM_OBJ_EXPLICIT_CENTER (mask px, may lie outside the map) */
    int32_t cls;                  /* OD_CLASS_IDX[obj_class] */
    int32_t flags;
    float peak;                   /* gaussian peak, 1.0 for ground truth */
    float track[2];               /* track_offset target (centertracker/processor.py:82-89) */
    float _pad;
} cvm_obj;

/* Ignore box, input-px numbers used as mask indices (processor.py:318-323). */
typedef struct cvm_box {
    double x, y, w, h;
} cvm_box;

/* Region of interest of one image (common/utils/image.py:9-19): inv_scale = (float)(1.0 / roi.scale). */
typedef struct cvm_roi {
    float inv_scale, off_left, off_top, _pad;
} cvm_roi;

const char* cvm_last_error(void);
int cvm_version(void);

/* ---- pre-processing front-end ------------------------------------------------------------------------------------ */
/* Raw labelled boxes -> render inputs, on the device: clip_to_img (processor.py:46-56) and the MIN_BOX_AREA filter
 * (processor.py:241-253) of ProcessImages.process for a whole batch; replaces the per-sample host loop of
 * BaseDataGenerator._process_batch (data/base_data_generator.py:29-48).  raw_boxes [n,4] fp64 (x, y, w, h in input px),
 * raw_cls [n] (OD_CLASS_IDX), raw_track [n,2] or NULL, raw_offsets [B+1] (image b owns [raw_offsets[b], raw_offsets[b+1])).
 * Outputs (capacity n each): objs / obj_offsets [B+1] in list order, boxes at or below the area limit become ignore
 * areas (ignore / ign_offsets [B+1]).  All pointers are device pointers. */
int cvm_prepare_objects(const double* raw_boxes, const int32_t* raw_cls, const float* raw_track, const int32_t* raw_offsets,
                        int B, double img_w, double img_h, double min_box_area, cvm_obj* objs, int32_t* obj_offsets,
                        cvm_box* ignore, int32_t* ign_offsets, void* stream);

/* ---- render -------------------------------------------------------------------------------------------------- */
/* objs[obj_offsets[b] .. obj_offsets[b+1]) are image b's objects in list order (order matters for the centre
 * scatter: last writer wins).  ignore / ign_offsets may be NULL.  y_true is [B,H,W,Ct], fully overwritten. */
int cvm_render_gt(const cvm_layout* L, const cvm_obj* objs, const int32_t* obj_offsets, const cvm_box* ignore,
                  const int32_t* ign_offsets, int B, float* y_true, void* stream);
/* cvm_render_gt plus per-object extra regression targets - l_shape (7 floats) and / or 3d_info (5 floats), computed on
 * the host from the 3D box (processor.py:69-115) - scattered to channels [extra_off, extra_off + extra_n) of the centre
 * pixel like the other targets (processor.py:296-299; last writer wins).  extra: device [n_obj][extra_stride] floats in
 * object order. */
int cvm_render_gt_extra(const cvm_layout* L, const cvm_obj* objs, const int32_t* obj_offsets, const cvm_box* ignore,
                        const int32_t* ign_offsets, int B, const float* extra, int extra_stride, int extra_off, int extra_n,
                        float* y_true, void* stream);
/* One plane [B,H,W,1], no weights, records use (cx, cy, w, h, peak) with CVM_OBJ_EXPLICIT_CENTER semantics. */
int cvm_render_prev_hm(const cvm_layout* L, const cvm_obj* objs, const int32_t* obj_offsets, int B, float* prev_hm,
                       void* stream);

/* Literal fill_heatmap drop-in (processor.py:18-22): max/min-combine `n_obj` records (explicit centres cx,cy; w,h;
 * peak) INTO an existing heat plane (element i at heat[i*heat_stride]) and, if not NULL, an existing weights plane. */
int cvm_fill_heatmap_inplace(const cvm_obj* objs, int n_obj, float* heat, int heat_stride, float* weights, int H, int W,
                             double R, double alpha, void* stream);

/* ---- loss ---------------------------------------------------------------------------------------------------- */
size_t cvm_loss_workspace_bytes(const cvm_layout* L, long long n_pixels);
/* n_pixels = B*H*W (the loss is batch-global, loss.py:50,113).  y_true_stride / y_pred_stride are the per-pixel
 * channel strides in floats (>= Ct / Cp; > when the CenterNet slice sits inside a wider multitask tensor,
 * multitask/loss.py:44-47).  partials[CVM_NPART] (device, fp64) = [P, N, n_pos, n_obj, field sums ...]: the vector
 * that is summed across GPUs before cvm_loss_finalize. */
int cvm_loss_fwd(const cvm_layout* L, const float* y_true, int y_true_stride, const float* y_pred, int y_pred_stride,
                 long long n_pixels, int use_weights, double* partials, void* ws, size_t ws_bytes, void* stream);
/* cvm_loss_fwd + cvm_loss_finalize in ONE launch, for callers that have nothing to sum across devices: the last block of
 * the forward kernel reduces the block partials (fixed order: bit-reproducible) and applies the finalise step right
 * away.  partials[CVM_NPART] is written as well (cvm_loss_bwd needs it); out as for cvm_loss_finalize. */
int cvm_loss_fwd_total(const cvm_layout* L, const float* y_true, int y_true_stride, const float* y_pred, int y_pred_stride,
                       long long n_pixels, int use_weights, double* partials, float* out, void* ws, size_t ws_bytes,
                       void* stream);
/* Data parallel: `gathered` [n_ranks][CVM_NPART] holds the partials of every rank (an all-gather of 128 bytes per rank).
 * They are summed in RANK ORDER - the result is bit-identical on every rank and to one GPU processing the shards one after
 * the other, whatever algorithm the collective used - written to `partials` (nullable; cvm_loss_bwd needs them) and
 * finalised into `out` (nullable) in the same launch. */
int cvm_loss_finalize_gathered(const cvm_layout* L, const double* gathered, int n_ranks, double* partials, float* out,
                               void* stream);
/* out[0] = total, out[1] = focal, out[2 + i] = field i (already normalised and post-transformed, unweighted). */
int cvm_loss_finalize(const cvm_layout* L, const double* partials, float* out, void* stream);
/* grad_pred [n_pixels, y_pred_stride-compatible: written with stride Cp] = upstream * d(total)/d(y_pred); needs the
 * globally reduced partials (n_pos, n_obj). */
int cvm_loss_bwd(const cvm_layout* L, const float* y_true, int y_true_stride, const float* y_pred, int y_pred_stride,
                 long long n_pixels, const double* partials, const float* upstream /* device scalar or NULL (=1) */,
                 float* grad_pred, void* stream);

/* cvm_loss_bwd through the layout-generic kernel only (the path of every layout without a compile-time instantiation); same
 * result within rounding - exposed so that callers / tests can compare the two. */
int cvm_loss_bwd_generic(const cvm_layout* L, const float* y_true, int y_true_stride, const float* y_pred, int y_pred_stride,
                         long long n_pixels, const double* partials, const float* upstream, float* grad_pred, void* stream);

/* ---- decode -------------------------------------------------------------------------------------------------- */
size_t cvm_decode_topk_workspace_bytes(const cvm_layout* L, int pred_stride, int B, int K);   /* 0 = unsupported shape */
/* y_pred [B,H,W,pred_stride] (pred_stride >= Cp).  Outputs (device): scores[B,K] f32, cls[B,K] i32,
 * flat[B,K] i64 (NHWC flat index (y*W+x)*hm+c), centers[B,K,2], boxes[B,K,4] (tlx,tly,w,h), track[B,K,2] (the tracking
 * OFFSET scaled to roi coordinates: centers + track = predicted centre in the previous frame, the query point of
 * cvm_track_associate; may be NULL).  Order: score desc, ties by lowest flat index (tf.nn.top_k).  rois: per image or NULL. */
int cvm_decode_topk(const cvm_layout* L, const float* y_pred, int pred_stride, int B, int K, const cvm_roi* rois,
                    float* scores, int32_t* cls, long long* flat, float* centers, float* boxes, float* track,
                    void* ws, size_t ws_bytes, void* stream);

/* cvm_decode_topk plus, in the SAME pass over y_pred, the per-pixel class pick of a semseg slice that lives in the same
 * tensor (multitask head: models/multitask/loss.py:21-32 slices, multitask/callbacks.py:107-110 + to_3channel's np.argmax,
 * common/utils/image.py:88: first maximum of channels [seg_off, seg_off + seg_n)).  seg_ids: u8 [B,H,W] class ids
 * (what cvm_semseg_argmax mode 0 writes).  The wide tensor is read once instead of twice. */
int cvm_decode_topk_semseg(const cvm_layout* L, const float* y_pred, int pred_stride, int B, int K, const cvm_roi* rois,
                           float* scores, int32_t* cls, long long* flat, float* centers, float* boxes, float* track,
                           int seg_off, int seg_n, unsigned char* seg_ids, void* ws, size_t ws_bytes, void* stream);

/* Process-wide setting: SMs the decode's persistent scan kernel leaves free (default 0).  A data-parallel caller that
 * issues the exchange of the loss partials right before the decode sets 1: the collective's kernel then runs beside the
 * scan instead of behind it (every scan CTA owns a whole SM's shared memory).  Affects the workspace size query too: set
 * it before sizing workspaces. */
int cvm_decode_set_spare_sms(int n);

/* The tiling cvm_decode_topk uses for a shape (introspection for tests and docs; no device work): out8 = pixels per granule,
 * granules per image, ring slots, ring mode (1: the 3x3 neighbours are read from shared memory, 0: from global memory),
 * halo granules, grid size, shared memory bytes, candidate buffer capacity. */
int cvm_decode_plan(const cvm_layout* L, int pred_stride, int B, int K, long long* out8);

/* Monitoring.  cvm_decode_topk predicts each image's K-th best score from the image before it (and, through the workspace,
 * from the previous call) so that only a few hundred pixels per image need the exact 3x3 test; the prediction is verified
 * per image and an image it does not hold for is recomputed exactly by a slower path.  Results never depend on it.  This
 * returns how many images (since the library was loaded, on the current device) took the slow path; it synchronises the
 * device.  -1 on error. */
long long cvm_decode_fallback_count(void);

/* Profile R: window x window first-argmax (window odd, 9 in the reference) + strict threshold, scan order.
 * counts[B] = number of objects found (may exceed max_out; only the first max_out in scan order are written). */
size_t cvm_decode_window9_workspace_bytes(const cvm_layout* L, int B);
int cvm_decode_window9(const cvm_layout* L, const float* y_pred, int pred_stride, int B, int window, float min_conf,
                       const cvm_roi* rois, int max_out, int32_t* counts, int32_t* cls, int32_t* pix /* y*W+x */,
                       float* scores, float* centers, float* boxes, void* ws, size_t ws_bytes, void* stream);

/* ---- semseg argmax -------------------------------------------------------------------------------------------- */
/* in [n_pixels, stride] floats, classes at [off, off+n_cls).  mode 0: write class ids u8 [n_pixels];
 * mode 1: write BGR u8 [n_pixels,3] with to_3channel semantics (lut_bgr [n_cls,3] device u8; threshold NaN = None). */
int cvm_semseg_argmax(const float* in, long long n_pixels, int stride, int off, int n_cls, int mode, int apply_softmax,
                      int use_weight, double threshold, const unsigned char* lut_bgr, unsigned char* out, void* stream);

/* ---- CenterTrack association ---------------------------------------------------------------------------------------- */
/* Greedy matching of the detections of one frame to the tracks of the previous frame, per image, as published with
 * CenterTrack (Zhou et al., "Tracking Objects as Points", ECCV 2020; src/lib/utils/tracker.py: Tracker.step +
 * greedy_assignment): detection i (in the given order = score order of cvm_decode_topk) looks for the closest still
 * unmatched previous centre to  centers[i] + track[i]  (squared distance d in fp32, x term + y term); a pair is invalid
 * if d > w*h of the previous box, d > w*h of the detection, or the classes differ; ties take the lowest previous index
 * (numpy argmin).  Detections with score < min_score (or NaN) are skipped.  Consumer of the `track` output of
 * cvm_decode_topk, whose targets are scattered by CenterTrackerProcess (models/centertracker/processor.py:77-89).
 *   centers/track [B,K,2], boxes [B,K,4] (tlx,tly,w,h), scores/cls [B,K]: outputs of cvm_decode_topk
 *   prev_centers [B,M,2], prev_sizes [B,M,2] (w,h), prev_cls [B,M], prev_count [B] (NULL = all M valid)
 *   match [B,K] int32: index of the matched previous track, -1 = none (a new track, or a skipped detection). */
int cvm_track_associate(const float* centers, const float* track, const float* boxes, const float* scores, const int32_t* cls,
                        int B, int K, float min_score, const float* prev_centers, const float* prev_sizes,
                        const int32_t* prev_cls, const int32_t* prev_count, int M, int32_t* match, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CVMHOT_H */
