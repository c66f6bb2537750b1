"""ctypes binding of libcvmhot.so (the C ABI declared in include/cvmhot.h).

There is deliberately NO fallback: if the shared library is missing or a call fails, an exception is raised.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# CVMHOT_LIB: load another build of the same library (tools/ build variants with -DCVM_EXPERIMENT / -DCVM_DECODE_STATS)
LIB_PATH = os.environ.get("CVMHOT_LIB") or os.path.join(_HERE, "lib", "libcvmhot.so")

CVM_MAX_FIELDS = 8
CVM_NPART = 16
KIND_MSE, KIND_MAE, KIND_MAPE, KIND_CE = 0, 1, 2, 3
POST_NONE, POST_ORIENT = 0, 1
OBJ_EXPLICIT_CENTER = 1
OBJ_NO_SCATTER = 2


class CvmLayout(C.Structure):
    _fields_ = [
        ("H", C.c_int32), ("W", C.c_int32), ("hm", C.c_int32), ("nb_classes", C.c_int32),
        ("Cp", C.c_int32), ("Ct", C.c_int32),
        ("off_class", C.c_int32), ("off_roff", C.c_int32), ("off_box", C.c_int32), ("off_track", C.c_int32),
        ("n_fields", C.c_int32),
        ("field_off", C.c_int32 * CVM_MAX_FIELDS), ("field_size", C.c_int32 * CVM_MAX_FIELDS),
        ("field_kind", C.c_int32 * CVM_MAX_FIELDS), ("field_post", C.c_int32 * CVM_MAX_FIELDS),
        ("field_weight", C.c_float * CVM_MAX_FIELDS),
        ("focal_a", C.c_float), ("focal_b", C.c_float),
        ("R", C.c_double), ("alpha", C.c_double),
    ]


class CvmError(RuntimeError):
    pass


_lib = None


def lib():
    """Load the shared library once. Raises CvmError when it has not been built (run __graft_entry__.build())."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise CvmError(f"{LIB_PATH} not found: build it with `make -C computer-vision-models_b200/csrc` "
                       "(there is no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    vp, i32, i64, f32, f64, sz = C.c_void_p, C.c_int, C.c_longlong, C.c_float, C.c_double, C.c_size_t
    LP = C.POINTER(CvmLayout)
    L.cvm_last_error.restype = C.c_char_p
    L.cvm_last_error.argtypes = []
    L.cvm_version.restype = i32
    L.cvm_version.argtypes = []
    L.cvm_prepare_objects.restype = i32
    L.cvm_prepare_objects.argtypes = [vp, vp, vp, vp, i32, f64, f64, f64, vp, vp, vp, vp, vp]
    L.cvm_render_gt.restype = i32
    L.cvm_render_gt.argtypes = [LP, vp, vp, vp, vp, i32, vp, vp]
    L.cvm_render_gt_extra.restype = i32
    L.cvm_render_gt_extra.argtypes = [LP, vp, vp, vp, vp, i32, vp, i32, i32, i32, vp, vp]
    L.cvm_render_prev_hm.restype = i32
    L.cvm_render_prev_hm.argtypes = [LP, vp, vp, i32, vp, vp]
    L.cvm_fill_heatmap_inplace.restype = i32
    L.cvm_fill_heatmap_inplace.argtypes = [vp, i32, vp, i32, vp, i32, i32, f64, f64, vp]
    L.cvm_loss_workspace_bytes.restype = sz
    L.cvm_loss_workspace_bytes.argtypes = [LP, i64]
    L.cvm_loss_fwd.restype = i32
    L.cvm_loss_fwd.argtypes = [LP, vp, i32, vp, i32, i64, i32, vp, vp, sz, vp]
    L.cvm_loss_fwd_total.restype = i32
    L.cvm_loss_fwd_total.argtypes = [LP, vp, i32, vp, i32, i64, i32, vp, vp, vp, sz, vp]
    L.cvm_loss_finalize_gathered.restype = i32
    L.cvm_loss_finalize_gathered.argtypes = [LP, vp, i32, vp, vp, vp]
    L.cvm_loss_finalize.restype = i32
    L.cvm_loss_finalize.argtypes = [LP, vp, vp, vp]
    L.cvm_loss_bwd.restype = i32
    L.cvm_loss_bwd.argtypes = [LP, vp, i32, vp, i32, i64, vp, vp, vp, vp]
    L.cvm_loss_bwd_generic.restype = i32
    L.cvm_loss_bwd_generic.argtypes = [LP, vp, i32, vp, i32, i64, vp, vp, vp, vp]
    L.cvm_decode_topk_workspace_bytes.restype = sz
    L.cvm_decode_topk_workspace_bytes.argtypes = [LP, i32, i32, i32]
    L.cvm_decode_topk.restype = i32
    L.cvm_decode_topk.argtypes = [LP, vp, i32, i32, i32, vp, vp, vp, vp, vp, vp, vp, vp, sz, vp]
    L.cvm_decode_topk_semseg.restype = i32
    L.cvm_decode_topk_semseg.argtypes = [LP, vp, i32, i32, i32, vp, vp, vp, vp, vp, vp, vp, i32, i32, vp, vp, sz, vp]
    L.cvm_decode_set_spare_sms.restype = i32
    L.cvm_decode_set_spare_sms.argtypes = [i32]
    L.cvm_decode_plan.restype = i32
    L.cvm_decode_plan.argtypes = [LP, i32, i32, i32, vp]
    L.cvm_decode_fallback_count.restype = i64
    L.cvm_decode_fallback_count.argtypes = []
    L.cvm_decode_window9_workspace_bytes.restype = sz
    L.cvm_decode_window9_workspace_bytes.argtypes = [LP, i32]
    L.cvm_decode_window9.restype = i32
    L.cvm_decode_window9.argtypes = [LP, vp, i32, i32, i32, f32, vp, i32, vp, vp, vp, vp, vp, vp, vp, sz, vp]
    L.cvm_semseg_argmax.restype = i32
    L.cvm_semseg_argmax.argtypes = [vp, i64, i32, i32, i32, i32, i32, i32, f64, vp, vp, vp]
    L.cvm_track_associate.restype = i32
    L.cvm_track_associate.argtypes = [vp, vp, vp, vp, vp, i32, i32, f32, vp, vp, vp, vp, i32, vp, vp]
    _lib = L
    return L


EXPORTS = [
    "cvm_last_error", "cvm_version", "cvm_prepare_objects", "cvm_render_gt", "cvm_render_gt_extra", "cvm_render_prev_hm", "cvm_fill_heatmap_inplace", "cvm_loss_workspace_bytes",
    "cvm_loss_fwd", "cvm_loss_fwd_total", "cvm_loss_finalize", "cvm_loss_finalize_gathered", "cvm_loss_bwd", "cvm_loss_bwd_generic", "cvm_decode_topk_workspace_bytes", "cvm_decode_topk", "cvm_decode_topk_semseg", "cvm_decode_fallback_count", "cvm_decode_plan", "cvm_decode_set_spare_sms",
    "cvm_decode_window9_workspace_bytes", "cvm_decode_window9", "cvm_semseg_argmax", "cvm_track_associate",
]


def check(rc, what):
    if rc != 0:
        raise CvmError(f"{what} failed (code {rc}): {lib().cvm_last_error().decode()}")
