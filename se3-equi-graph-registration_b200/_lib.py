"""Loader / builder of the C-ABI shared library `libegspr_b200.so` (include/egspr_b200.h).

The library is built IN-TREE with nvcc for sm_100a and loaded with ctypes.  There is no fallback:
if the library is missing or a symbol is absent the import of any op fails loudly.
"""
import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_CSRC = os.path.join(_HERE, "csrc")
LIB_PATH = os.path.join(_HERE, "libegspr_b200.so")
SOURCES = ["knn.cu", "csr.cu", "egnn_layer.cu", "egnn_edge_ts.cu", "egnn_node_ts.cu", "head.cu", "feature_match.cu",
           "egnn_backward.cu", "egnn_edge_bwd_tc.cu"]
HEADERS = ["egspr_common.cuh", "egnn_layer.cuh", "tcgen05.cuh", "egnn_backward_math.cuh", "egnn_backward.cuh"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]

_p = ctypes.c_void_p
_i = ctypes.c_int
_l = ctypes.c_int64
_f = ctypes.c_float
_z = ctypes.c_size_t

# name -> (restype, argtypes); must list every symbol include/egspr_b200.h declares
SIGNATURES = {
    "egspr_version": (_i, []),
    "egspr_error_string": (ctypes.c_char_p, [_i]),
    "egspr_knn_workspace_bytes": (_z, [_i, _i]),
    "egspr_knn_build": (_i, [_p, _i, _i, _i, _p, _p, _z, _p]),
    "egspr_nbr_to_edges": (_i, [_p, _i, _i, _i, _p, _p]),
    "egspr_csr_workspace_bytes": (_z, [_l, _l]),
    "egspr_csr_from_nbr": (_i, [_p, _i, _i, _i, _p, _p, _p, _p, _p, _z, _p, _p]),
    "egspr_csr_from_edges": (_i, [_p, _i, _i, _l, _p, _p, _p, _p, _p, _z, _p, _p]),
    "egspr_segment_sum": (_i, [_p, _i, _p, _p, _l, _p, _p]),
    "egspr_node_embed": (_i, [_p, _p, _l, _p, _p, _p, _p, _p, _p, _p]),
    "egspr_egcl_forward": (_i, [_p, _p, _p, _p, _p, _p, _p, _p, _p, _f, _l, _l, _i, _p, _p, _p,
                                _p, _p, _p, _p, _p, _p, _i, _p]),
    "egspr_kabsch": (_i, [_p, _p, _p, _p, _i, _i, _p, _p, _p, _p]),
    "egspr_head_eval": (_i, [_p] * 11 + [_i, _i, _i] + [_p] * 6),
    "egspr_head_eval_workspace_bytes": (_z, [_i]),
    "egspr_head_eval_ws": (_i, [_p] * 11 + [_i, _i, _i] + [_p] * 6 + [_z, _p]),
    "egspr_head_train": (_i, [_p] * 6 + [_i, _i] + [_p] * 7),
    "egspr_pose_metrics": (_i, [_p] * 5 + [_i, _i, ctypes.c_double, _p, _p]),
    "egspr_feature_nn": (_i, [_p, _i, _p, _i, _p, _z, _p, _p, _p]),
    "egspr_egcl_backward_workspace_bytes": (_z, [_l, _l]),
    "egspr_csr_edge_positions": (_i, [_p, _p, _p, _i, _l, _l, _p, _p, _p]),
    "egspr_egcl_backward": (_i, [_p] * 11 + [_p, _f, _l, _l, _i] + [_p] * 7 + [_z, _p]),
    "egspr_linear32_forward": (_i, [_p, _l, _p, _p, _p]),
    "egspr_linear32_backward": (_i, [_p, _p, _l, _p, _p, _p, _p]),
    "egspr_head_train_backward": (_i, [_p] * 8 + [_i, _i] + [_p] * 5),
    "egspr_pose_loss": (_i, [_p, _p, _p, _i, _p, _p, _p, _p, _p]),
    "egspr_train_loss_forward": (_i, [_p] * 7 + [_i, _i, _i] + [_p] * 6),
    "egspr_train_loss_finalize": (_i, [_p] * 4 + [_i, _i, _i] + [_p] * 3 + [_f] + [_p] * 5),
    "egspr_head_train_loss_backward": (_i, [_p] * 11 + [_i, _i, _i] + [_p] * 6),
}


def sources():
    return [os.path.join(_CSRC, s) for s in SOURCES]


def needs_build():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = sources() + [os.path.join(_CSRC, h) for h in HEADERS] + [os.path.join(_HERE, "..", "include", "egspr_b200.h")]
    return any(os.path.exists(d) and os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    """Compile every CUDA source for sm_100a into the in-tree shared library."""
    if not force and not needs_build():
        return LIB_PATH
    cmd = ["nvcc"] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB_PATH] + sources()
    subprocess.check_call(cmd, cwd=_CSRC)
    return LIB_PATH


_LIB = None


def lib():
    """The loaded library with typed entry points.  Raises if it was never built -- the product
    path has no CPU or PyTorch fallback."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a).  There is no CPU fallback for the registration hot path.")
        handle = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)       # AttributeError if the symbol is not exported
            fn.restype = res
            fn.argtypes = args
        _LIB = handle
    return _LIB


class EgsprError(RuntimeError):
    pass


def check(code, what):
    if code != 0:
        msg = lib().egspr_error_string(code).decode()
        raise EgsprError(f"{what} failed: {msg} (code {code})")
