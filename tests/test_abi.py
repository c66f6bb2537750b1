"""CPU: the C-ABI shared library loads without a GPU and exports every symbol include/cvmhot.h declares; argument
validation and workspace queries work on the host; the product path fails loudly without CUDA tensors."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "cvmhot.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(cvm_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported():
    from cvmhot import _lib
    lib = _lib.lib()
    declared = _declared_symbols()
    assert len(declared) >= 14
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in cvmhot.h but not exported"
    assert sorted(_lib.EXPORTS) == declared
    assert lib.cvm_version() >= 100


def test_struct_sizes_match_header():
    from cvmhot import _lib, ops
    assert ops.OBJ_DTYPE.itemsize == 64 and ops.BOX_DTYPE.itemsize == 32 and ops.ROI_DTYPE.itemsize == 16
    # 11 int32 + 4*8 int32 + 8 float + 2 float + 2 double
    assert ctypes.sizeof(_lib.CvmLayout) == 11 * 4 + 4 * 8 * 4 + 8 * 4 + 2 * 4 + 2 * 8 + 4   # +4 padding before the doubles


def test_host_side_validation_and_queries():
    from cvmhot import _lib
    from cvmhot.layout import layout_from_params
    from cvmhot.models.centernet import CenternetParams
    lib = _lib.lib()
    p = CenternetParams(10, per_class_heatmap=True)
    p.INPUT_HEIGHT, p.INPUT_WIDTH = 256, 768
    L = layout_from_params(p)
    assert (L.H, L.W, L.hm, L.Cp, L.Ct) == (128, 384, 10, 14, 15)
    s = L.c_struct()
    assert lib.cvm_loss_workspace_bytes(ctypes.byref(s), 128 * 384 * 8) > 0
    assert lib.cvm_decode_topk_workspace_bytes(ctypes.byref(s), 14, 256, 100) > 0
    assert lib.cvm_decode_window9_workspace_bytes(ctypes.byref(s), 4) >= 4 * 128 * 384 * 5
    # bad arguments are rejected on the host, with a message, before any launch
    rc = lib.cvm_loss_fwd(ctypes.byref(s), None, 15, None, 14, 10, 1, None, None, 0, None)
    assert rc == -1 and b"NULL" in lib.cvm_last_error()
    assert lib.cvm_decode_topk_workspace_bytes(ctypes.byref(s), 14, 256, 5000) == 0          # K too large
    s.hm = 200
    assert lib.cvm_loss_fwd(ctypes.byref(s), ctypes.c_void_p(16), 15, ctypes.c_void_p(16), 14, 10, 1,
                            ctypes.c_void_p(16), ctypes.c_void_p(16), 1 << 20, None) == -1


def test_no_cpu_fallback():
    import torch
    from cvmhot import _lib, ops
    from cvmhot.layout import layout_from_params
    from cvmhot.models.centernet import CenternetParams
    L = layout_from_params(CenternetParams(3))
    with pytest.raises(_lib.CvmError):
        ops.loss_partials(L, torch.zeros(1, L.H, L.W, L.Ct), torch.zeros(1, L.H, L.W, L.Cp))
    with pytest.raises(_lib.CvmError):
        ops.decode_topk(L, torch.zeros(1, L.H, L.W, L.Cp))


def test_product_does_not_import_oracle():
    """The oracle is test infrastructure: nothing under the product package may reference it."""
    pkg = os.path.join(ROOT, "computer-vision-models_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in txt, os.path.join(dirpath, f)
