"""ctypes binding of libsdtf.so (include/sdtf.h) with DLPack tensor exchange.

There is deliberately no fallback here: if the shared library is missing or no sm_100 GPU is present, loading /
engine creation raises.  Tensors are handed to C as `DLManagedTensor*` taken from a genuine "dltensor" PyCapsule
(`ndarray.__dlpack__()` / `torch.utils.dlpack.to_dlpack`), borrowed for the duration of the call.
"""
from __future__ import annotations

import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsdtf.so")

EXPORTS = [
    "sdtf_create", "sdtf_destroy", "sdtf_last_error", "sdtf_version", "sdtf_load_tensor", "sdtf_finalize_weights",
    "sdtf_unet_forward", "sdtf_controlnet_forward", "sdtf_hintnet_forward", "sdtf_vae_decode", "sdtf_vae_encode", "sdtf_text_encode",
    "sdtf_text_embed", "sdtf_text_encode_embedded",
    "sdtf_cfg_sched_step", "sdtf_to_uint8", "sdtf_denoise", "sdtf_get_timings", "sdtf_bench_conv", "sdtf_bench_attention",
    "sdtf_test_attention", "sdtf_test_norm", "sdtf_trace_begin", "sdtf_trace_end", "sdtf_comm_unique_id", "sdtf_comm_init", "sdtf_comm_destroy",
]


class StepCoef(ctypes.Structure):
    _fields_ = [("guidance", ctypes.c_float), ("rescale", ctypes.c_float), ("ca", ctypes.c_float),
                ("cb", ctypes.c_float), ("cn", ctypes.c_float), ("sig_t", ctypes.c_float), ("noi_t", ctypes.c_float)]


class DenoiseDesc(ctypes.Structure):
    _fields_ = [("n_steps", ctypes.c_int32), ("use_cuda_graph", ctypes.c_int32), ("decode", ctypes.c_int32),
                ("cfg_split", ctypes.c_int32),
                ("latent0", ctypes.c_void_p), ("context", ctypes.c_void_p), ("uncond_context", ctypes.c_void_p),
                ("t_emb", ctypes.c_void_p), ("coefs", ctypes.POINTER(StepCoef)), ("step_noise", ctypes.c_void_p),
                ("mask", ctypes.c_void_p), ("init_latent", ctypes.c_void_p), ("init_noise", ctypes.c_void_p),
                ("hint_image", ctypes.c_void_p), ("blend_image", ctypes.c_void_p), ("blend_mask", ctypes.c_void_p),
                ("out_images", ctypes.c_void_p), ("out_latent", ctypes.c_void_p),
                ("on_step", ctypes.c_void_p), ("on_step_user", ctypes.c_void_p)]


ON_STEP = ctypes.CFUNCTYPE(None, ctypes.c_int32, ctypes.c_void_p)  # void (*on_step)(int32_t iteration, void* user)


class Timings(ctypes.Structure):
    _fields_ = [("loop_ms", ctypes.c_float), ("decode_ms", ctypes.c_float), ("total_ms", ctypes.c_float),
                ("kernel_launches", ctypes.c_int32)]


class TraceSummary(ctypes.Structure):
    _fields_ = [("launches", ctypes.c_int64 * 4), ("us", ctypes.c_double * 4), ("flop", ctypes.c_double * 4),
                ("bytes", ctypes.c_double * 4)]


_lib = None


def load():
    """Load libsdtf.so (building it first if the sources are newer and nvcc is around)."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("SDTF_LIB") or LIB_PATH  # SDTF_LIB: an alternative build of the same library (A/B measurements)
    if path == LIB_PATH and not os.path.exists(LIB_PATH):
        from . import build
        build.build_lib()
    if not os.path.exists(path):
        raise RuntimeError(f"{path} is missing: build it with `python -m minsdtf_b200.build` — there is no fallback path")
    lib = ctypes.CDLL(path)
    vp, i32 = ctypes.c_void_p, ctypes.c_int32
    lib.sdtf_create.argtypes = [i32, ctypes.POINTER(vp)]
    lib.sdtf_destroy.argtypes = [vp]
    lib.sdtf_destroy.restype = None
    lib.sdtf_last_error.argtypes = [vp]
    lib.sdtf_last_error.restype = ctypes.c_char_p
    lib.sdtf_version.restype = ctypes.c_char_p
    lib.sdtf_load_tensor.argtypes = [vp, ctypes.c_char_p, vp]
    lib.sdtf_finalize_weights.argtypes = [vp, ctypes.c_char_p]
    lib.sdtf_unet_forward.argtypes = [vp, vp, vp, vp, ctypes.POINTER(vp), vp]
    lib.sdtf_controlnet_forward.argtypes = [vp, vp, vp, vp, vp, ctypes.POINTER(vp)]
    lib.sdtf_hintnet_forward.argtypes = [vp, vp, vp]
    lib.sdtf_vae_decode.argtypes = [vp, vp, vp]
    lib.sdtf_vae_encode.argtypes = [vp, vp, vp]
    lib.sdtf_text_encode.argtypes = [vp, vp, i32, vp]
    lib.sdtf_text_embed.argtypes = [vp, vp, vp, vp]
    lib.sdtf_text_encode_embedded.argtypes = [vp, vp, i32, vp]
    lib.sdtf_cfg_sched_step.argtypes = [vp, vp, vp, vp, ctypes.POINTER(StepCoef), vp, vp, vp, vp, vp]
    lib.sdtf_to_uint8.argtypes = [vp, vp, vp, vp, vp]
    lib.sdtf_denoise.argtypes = [vp, ctypes.POINTER(DenoiseDesc)]
    lib.sdtf_get_timings.argtypes = [vp, ctypes.POINTER(Timings)]
    lib.sdtf_bench_conv.argtypes = [vp, i32, i32, i32, i32, i32, i32, ctypes.POINTER(ctypes.c_float)]
    lib.sdtf_bench_attention.argtypes = [vp, i32, i32, i32, i32, i32, i32, i32, ctypes.POINTER(ctypes.c_float)]
    lib.sdtf_comm_unique_id.argtypes = [ctypes.c_char_p, vp]
    lib.sdtf_comm_init.argtypes = [vp, ctypes.c_char_p, vp, i32, i32]
    lib.sdtf_comm_destroy.argtypes = [vp]
    lib.sdtf_trace_begin.argtypes = [vp]
    lib.sdtf_trace_end.argtypes = [vp, ctypes.POINTER(TraceSummary)]
    lib.sdtf_test_attention.argtypes = [vp, vp, vp, vp, i32, vp]
    lib.sdtf_test_norm.argtypes = [vp, vp, vp, vp, i32, vp]
    for name in EXPORTS:
        fn = getattr(lib, name)
        if fn.restype is ctypes.c_int:  # default
            fn.restype = ctypes.c_int
    _lib = lib
    return lib


_PyCapsule_GetPointer = ctypes.pythonapi.PyCapsule_GetPointer
_PyCapsule_GetPointer.restype = ctypes.c_void_p
_PyCapsule_GetPointer.argtypes = [ctypes.py_object, ctypes.c_char_p]


def is_torch(x) -> bool:
    return type(x).__module__.startswith("torch")


class DL:
    """Keeps the DLPack capsules of one call alive and hands out their DLManagedTensor* addresses."""

    def __init__(self):
        self._keep = []

    def __call__(self, x):
        if x is None:
            return None
        if is_torch(x):
            import torch.utils.dlpack
            x = x.detach()
            if not x.is_contiguous():
                x = x.contiguous()
            cap = torch.utils.dlpack.to_dlpack(x)
        else:
            x = np.asarray(x)
            if not (x.flags.c_contiguous and x.flags.writeable and x.flags.aligned):
                x = np.array(x, order="C", copy=True)
            cap = x.__dlpack__()
        self._keep.append((x, cap))
        return _PyCapsule_GetPointer(cap, b"dltensor")

    def array(self, xs):
        ptrs = [self(x) for x in xs]
        return (ctypes.c_void_p * len(ptrs))(*ptrs)
