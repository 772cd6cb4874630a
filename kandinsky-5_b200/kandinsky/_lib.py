"""ctypes binding of libk5.so (include/k5.h).  There is no fallback: if the CUDA library is missing the
import fails loudly, and every call that returns a non-zero code raises."""
import ctypes
import os
from ctypes import POINTER, Structure, c_char_p, c_float, c_int, c_int32, c_int64, c_uint8, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
# K5_LIB_PATH selects another build of the same library (A/B of compile-time tuning knobs); there is still no fallback
LIB_PATH = os.environ.get("K5_LIB_PATH") or os.path.join(os.path.dirname(_HERE), "libk5.so")

K5_OK, K5_ERR_INVALID, K5_ERR_CUDA, K5_ERR_STATE, K5_ERR_UNSUPPORTED = 0, 1, 2, 3, 4
EPI_STORE, EPI_GELU, EPI_GATE, EPI_HEADS, EPI_F32 = 0, 1, 2, 3, 4
DIST_HANDLE_BYTES = 192
DTYPE_F32, DTYPE_BF16, DTYPE_F16 = 0, 1, 2


class K5Config(Structure):
    _fields_ = [
        ("in_visual_dim", c_int32), ("out_visual_dim", c_int32), ("time_dim", c_int32), ("patch_size", c_int32 * 3),
        ("model_dim", c_int32), ("ff_dim", c_int32), ("num_text_blocks", c_int32), ("num_visual_blocks", c_int32),
        ("axes_dims", c_int32 * 3), ("visual_cond", c_int32), ("in_text_dim", c_int32), ("in_text_dim2", c_int32),
        ("max_tokens", c_int32), ("max_text_tokens", c_int32),
    ]


class K5VaeConfig(Structure):
    _fields_ = [("block_out_channels", c_int32 * 4), ("latent_channels", c_int32), ("out_channels", c_int32),
                ("max_tile_frames", c_int32), ("max_height", c_int32), ("max_width", c_int32)]


class K5Sparse(Structure):
    _fields_ = [("P", c_float), ("wT", c_int32), ("wH", c_int32), ("wW", c_int32), ("add_sta", c_int32)]


# name -> (restype, argtypes); every symbol include/k5.h declares
SIGNATURES = {
    "k5_last_error": (c_char_p, []),
    "k5_version": (c_int, []),
    "k5_engine_create": (c_int, [POINTER(K5Config), POINTER(c_void_p)]),
    "k5_engine_destroy": (None, [c_void_p]),
    "k5_engine_load_tensor": (c_int, [c_void_p, c_char_p, c_void_p, c_int, POINTER(c_int64), c_int]),
    "k5_engine_finalize": (c_int, [c_void_p]),
    "k5_engine_set_grid": (c_int, [c_void_p, c_int, c_int, c_int, POINTER(c_int32), POINTER(c_int32), POINTER(c_int32),
                                   POINTER(c_float), c_int]),
    "k5_dit_forward": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_int, POINTER(c_int32), c_void_p, c_float,
                               POINTER(K5Sparse), c_void_p, c_void_p]),
    "k5_dit_forward_magcache": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_int, POINTER(c_int32), c_void_p, c_float,
                                        POINTER(K5Sparse), c_void_p, c_int, c_int, c_void_p]),
    "k5_sample": (c_int, [c_void_p, c_void_p, c_int, c_float, c_float, c_void_p, c_int, c_void_p, c_void_p, c_int,
                          c_void_p, POINTER(K5Sparse), c_void_p]),
    "k5_sample_magcache": (c_int, [c_void_p, c_void_p, c_int, c_float, c_float, c_void_p, c_int, c_void_p, c_void_p, c_int,
                                   c_void_p, POINTER(K5Sparse), c_void_p, c_void_p]),
    "k5_engine_attention_timing": (c_int, [c_void_p, c_int, POINTER(ctypes.c_double), POINTER(c_int64)]),
    "k5_dist_export": (c_int, [c_void_p, c_void_p]),
    "k5_dist_init": (c_int, [c_void_p, c_int, c_int, c_void_p]),
    "k5_dist_barrier": (c_int, [c_void_p, c_void_p]),
    "k5_dist_local_frames": (c_int, [c_void_p, POINTER(c_int), POINTER(c_int)]),
    "k5_dist_mode": (c_int, [c_void_p]),
    "k5_vae_create": (c_int, [POINTER(K5VaeConfig), POINTER(c_void_p)]),
    "k5_vae_destroy": (None, [c_void_p]),
    "k5_vae_load_tensor": (c_int, [c_void_p, c_char_p, c_void_p, c_int, POINTER(c_int64), c_int]),
    "k5_vae_finalize": (c_int, [c_void_p]),
    "k5_vae_decode": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "k5_conv3d_causal": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_int, c_void_p, c_void_p, c_void_p,
                                 c_void_p, c_void_p]),
    "k5_launch_count": (c_int64, [c_int]),
    "k5_last_sparse_density": (c_float, [c_void_p]),
    "k5_gemm_bf16": (c_int, [c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_int, c_void_p,
                             c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
    "k5_attention": (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int,
                             c_float, c_void_p, c_void_p, c_void_p]),
    "k5_attention_bounded": (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_int, c_void_p, c_int, c_int, c_int,
                                     c_int, c_float, c_void_p, c_void_p, c_float, c_void_p]),
    "k5_attention_bounded_split": (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_int, c_void_p, c_int, c_int, c_int,
                                           c_int, c_float, c_float, c_int, c_int, c_void_p, c_void_p]),
    "k5_debug_attn_trace": (c_int, [c_void_p]),
    "k5_ln_rows": (c_int, [c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_int, c_float, c_void_p]),
    "k5_nabla_select": (c_int, [c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_float, c_void_p, c_void_p, c_void_p,
                                c_void_p, c_void_p]),
    "k5_sta_mask": (c_int, [c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
}

_lib = None


def lib():
    """Load libk5.so once.  Raises ImportError with build instructions when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a).  There is no CPU or PyTorch fallback for the DiT hot path.")
        handle = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(rc):
    """Map C-ABI error codes to the exceptions the reference raises (ValueError for bad arguments)."""
    if rc == K5_OK:
        return
    msg = lib().k5_last_error().decode("utf-8", "replace")
    if rc == K5_ERR_INVALID:
        raise ValueError(msg)
    if rc == K5_ERR_UNSUPPORTED:
        raise NotImplementedError(msg)
    raise RuntimeError(f"libk5 error {rc}: {msg}")


def ptr(t):
    """data_ptr of a tensor (or None) as c_void_p."""
    return None if t is None else c_void_p(t.data_ptr())


def stream_ptr():
    import torch

    return c_void_p(torch.cuda.current_stream().cuda_stream)
