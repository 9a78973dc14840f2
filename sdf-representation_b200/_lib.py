"""ctypes binding of libsdfnet.so (include/sdfnet.h).  There is no fallback: a missing library is an error."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# SDFNET_LIB selects another build of the same sources (tools/build_prof.sh: the cycle-counter measurement build)
LIB_PATH = os.environ.get("SDFNET_LIB") or os.path.join(_HERE, "libsdfnet.so")

c_void_p = ctypes.c_void_p
c_int32 = ctypes.c_int32
c_int64 = ctypes.c_int64
c_float = ctypes.c_float
c_size_t = ctypes.c_size_t


class SdfnetDesc(ctypes.Structure):
    _fields_ = [
        ("d_in", c_int32),
        ("hidden", c_int32),
        ("num_hidden", c_int32),
        ("skip_layer", c_int32),
        ("beta", c_float),
        ("dtype", c_int32),
        ("precision", c_int32),
        ("device", c_int32),
    ]


class SdfnetIgrCfg(ctypes.Structure):
    _fields_ = [
        ("delta", c_float),
        ("tau", c_float),
        ("lambda_g", c_float),
        ("regularizer_threshold", c_float),
        ("data_term", c_int32),
        ("mean_t", c_float),
        ("n_global", c_int64),
        ("mean_t_dev", c_void_p),
        ("label_stats_dev", c_void_p),
    ]


# name -> (restype, argtypes); every name here is declared in include/sdfnet.h
SIGNATURES = {
    "sdfnet_last_error": (ctypes.c_char_p, []),
    "sdfnet_create": (c_int32, [ctypes.POINTER(SdfnetDesc), ctypes.POINTER(c_void_p)]),
    "sdfnet_destroy": (c_int32, [c_void_p]),
    "sdfnet_pack_weights": (c_int32, [c_void_p, ctypes.POINTER(c_void_p), ctypes.POINTER(c_void_p), c_void_p]),
    "sdfnet_workspace_bytes": (c_size_t, [c_void_p, c_int64, c_int32]),
    "sdfnet_forward": (c_int32, [c_void_p, c_void_p, c_int64, c_void_p, c_void_p]),
    "sdfnet_forward_grad": (c_int32, [c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "sdfnet_grid_eval": (c_int32, [c_void_p, c_void_p, c_int32, c_int64, c_int64, c_void_p, c_void_p]),
    "sdfnet_octree_level": (c_int32, [c_void_p, c_void_p, c_int64, c_int32, c_float, c_void_p, c_void_p, c_void_p]),
    "sdfnet_octree_levels": (
        c_int32,
        [c_void_p, c_int32, c_int32, c_int32, c_float, c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p],
    ),
    "sdfnet_backward": (c_int32, [c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "sdfnet_zero_grads": (c_int32, [c_void_p, c_void_p]),
    "sdfnet_read_grads": (c_int32, [c_void_p, c_void_p, c_float, c_void_p]),
    "sdfnet_num_params": (c_int64, [c_void_p]),
    "sdfnet_train_step_igr": (
        c_int32,
        [c_void_p, c_void_p, c_void_p, c_int64, ctypes.POINTER(SdfnetIgrCfg), c_void_p, c_void_p, c_size_t, c_void_p],
    ),
    "sdfnet_train_step_pcd": (
        c_int32,
        [c_void_p, c_void_p, c_int64, c_int64, c_void_p, c_int64, c_int64, c_float, c_void_p, c_void_p, c_size_t, c_void_p],
    ),
    "sdfnet_wsl2_label_stats": (c_int32, [c_void_p, c_int64, c_float, c_float, c_void_p, c_void_p]),
    "sdfnet_train_step_wsl2": (
        c_int32,
        [c_void_p, c_void_p, c_void_p, c_int64, c_float, c_float, c_int32, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p],
    ),
    "sdfnet_adam_step": (
        c_int32,
        [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_float, c_float, c_float, c_float, c_int32, c_void_p],
    ),
    "sdfnet_fused_adam_step": (
        c_int32,
        [c_void_p, ctypes.POINTER(c_void_p), ctypes.POINTER(c_void_p), c_void_p, c_void_p, c_float, c_float, c_float, c_float, c_int32,
         c_float, c_void_p, c_void_p, c_void_p],
    ),
    "sdfnet_set_deterministic": (c_int32, [c_void_p, c_int32]),
    "sdfnet_label_stats": (c_int32, [c_void_p, c_int64, c_float, c_void_p, c_void_p]),
    "sdfnet_grad_buffer": (c_int32, [c_void_p, ctypes.POINTER(c_void_p), ctypes.POINTER(c_int64)]),
    "sdfnet_loss_value": (c_int32, [c_void_p, c_int32, c_float, c_float, c_int32, c_int64, c_int64, c_void_p, c_void_p]),
    "sdfnet_launch_count": (c_int64, [c_void_p]),
    "sdfnet_profile_enable": (c_int32, [c_void_p, c_int32]),
    "sdfnet_profile_read": (c_int32, [c_void_p, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(c_int64)]),
    "sdfnet_mc_classify": (c_int32, [c_void_p, c_int32, c_int32, c_int32, c_void_p, c_void_p, c_void_p, c_void_p]),
    "sdfnet_mc_compact": (c_int32, [c_void_p, c_int64, c_void_p, c_void_p, c_void_p]),
    "sdfnet_mc_edge_vertices": (c_int32, [c_void_p, c_int32, c_int32, c_int32, c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p]),
    "sdfnet_mc_triangles": (c_int32, [c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p]),
    "sdfnet_octree_refine": (c_int32, [c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p]),
}

_lib = None


def load():
    """Load libsdfnet.so and bind every entry point.  Raises if the extension has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build the CUDA extension first (python -c 'import __graft_entry__ as g; g.build()'). "
            "There is no CPU or PyTorch fallback for the hot path."
        )
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc, what=""):
    if rc != 0:
        msg = load().sdfnet_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"{what or 'sdfnet'} failed (code {rc}): {msg}")
