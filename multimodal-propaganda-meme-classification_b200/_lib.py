"""ctypes binding of libb200mm.so, the C-ABI CUDA library (declared in include/b200mm.h).

There is no CPU or PyTorch fallback: if the library cannot be loaded, or an entry point returns a
non-zero status, the caller gets an exception.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_int, c_longlong, c_void_p, c_float, c_ulonglong

from . import _build

_lib = None

c_ptr = c_void_p

# name -> argtypes (restype is always int: 0 = ok, >0 = cudaError_t, <0 = b200mm error)
_SIGNATURES: dict[str, list] = {}


def declare(name: str, argtypes: list) -> None:
    _SIGNATURES[name] = argtypes


class B200MMError(RuntimeError):
    pass


_ERRORS = {-1: "bad argument (shape/alignment contract)", -2: "CUDA driver entry point unavailable",
           -3: "tensor map rejected by the driver", -4: "device is not sm_100 (B200)",
           -10: "JPEG coding not handled by the split decoder", -11: "not a JPEG, or damaged / truncated"}


def load(build_if_needed: bool = True) -> ctypes.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    path = _build.LIB_PATH
    if build_if_needed and _build.needs_build():
        # The sources do not match the stamp of the shipped binary.  With a compiler present the library is rebuilt and
        # ANY compile / link error propagates (a stale .so must never stand in for kernels that no longer build).
        # Only a box without nvcc may load the binary that travelled with the repo -- loudly.
        if _build.have_nvcc():
            _build.build()
        elif os.path.exists(path):
            import warnings
            warnings.warn("b200mm: csrc/ differs from the stamp of the prebuilt libb200mm.so and nvcc is not available "
                          "here; loading the prebuilt binary (set B200MM_STRICT_STAMP=1 to make this an error)",
                          RuntimeWarning, stacklevel=2)
            if os.environ.get("B200MM_STRICT_STAMP", "0") == "1":
                raise B200MMError("libb200mm.so is stale (source digest differs from build/stamp) and cannot be rebuilt")
        else:
            raise B200MMError("libb200mm.so missing and nvcc not found: it cannot be built (there is no fallback path)")
    if not os.path.exists(path):
        raise B200MMError(f"{path} not found: run `python -c 'import __graft_entry__ as g; g.build()'`")
    lib = ctypes.CDLL(path)
    for name, argtypes in _SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here = header/library mismatch: fail loudly
        fn.argtypes = argtypes
        fn.restype = c_int
    _lib = lib
    return lib


# kernels launched through this module since the counter was last reset (bench.py's gpu_launches)
LAUNCHES = [0]
_KERNELS_PER_CALL = {"b200mm_batchnorm_fwd": 2, "b200mm_jpeg_parse": 0, "b200mm_jpeg_entropy_decode": 0,
                     "b200mm_jpeg_reconstruct": 2, "b200mm_augment_pil": 2, "b200mm_jpeg_entropy_decode_sparse": 0, "b200mm_jpeg_reconstruct_sparse": 2, "b200mm_augment_jitter_rotate": 2, "b200mm_batchnorm_bwd": 2, "b200mm_version": 0, "b200mm_num_sms": 0, "b200mm_tune": 0,
                     "b200mm_set_step_salt_ptr": 0}


# optional per-call CUDA-event profile of a real (pipelined, warm-L2) run: PROFILE = [] enables it;
# entries are (name, key, start_event, end_event).  See scripts/profile_step.py.
PROFILE = None


def call(name: str, *args, key=None) -> None:
    LAUNCHES[0] += _KERNELS_PER_CALL.get(name, 1)
    if PROFILE is not None:
        import torch
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = getattr(load(), name)(*args)
        e1.record()
        PROFILE.append((name, key, e0, e1))
    else:
        rc = getattr(load(), name)(*args)
    if rc != 0:
        if rc > 0:
            raise B200MMError(f"{name}: CUDA error {rc}")
        raise B200MMError(f"{name}: {_ERRORS.get(rc, rc)}")


def exported_symbols() -> list[str]:
    return sorted(_SIGNATURES)


# ------------------------------------------------------------------ signatures (mirror include/b200mm.h)
declare("b200mm_version", [])
declare("b200mm_num_sms", [])
declare("b200mm_tune", [c_int, c_int])
declare("b200mm_gemm_bf16_maskres", [c_ptr, c_int, c_longlong, c_ptr, c_int, c_longlong, c_int, c_int, c_int, c_ptr,
                                     c_longlong, c_ptr, c_longlong, c_ptr, c_longlong, c_ptr])
declare("b200mm_gemm_bf16", [c_ptr, c_int, c_longlong, c_ptr, c_int, c_longlong, c_int, c_int, c_int, c_int,
                             c_ptr, c_ptr, c_longlong, c_ptr, c_longlong, c_ptr, c_longlong, c_ptr, c_longlong,
                             c_int, c_int, c_float, c_ulonglong, c_ptr, c_ptr])
declare("b200mm_conv_fwd", [c_ptr, c_int, c_int, c_int, c_int, c_ptr, c_int, c_int, c_int, c_int, c_int, c_ptr, c_ptr,
                            c_longlong, c_ptr, c_longlong, c_ptr, c_ptr])
declare("b200mm_conv_wgrad", [c_ptr, c_longlong, c_ptr, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_ptr,
                              c_int, c_ptr])
declare("b200mm_conv_weight_rotate", [c_ptr, c_ptr, c_int, c_int, c_int, c_ptr])
declare("b200mm_conv_weight_rotate_multi", [c_ptr, c_int, c_ptr])
declare("b200mm_preprocess_u8", [c_ptr, c_ptr, c_ptr, c_int, c_int, c_int, c_ptr, c_ptr, c_ptr, c_ptr])
declare("b200mm_preprocess_u8_packed", [c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_int, c_int, c_int, c_int, c_ptr, c_ptr, c_ptr,
                                        c_ptr])
declare("b200mm_preprocess_u8_packed_pil", [c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_int, c_int, c_int, c_int, c_ptr, c_ptr, c_ptr,
                                        c_ptr])
declare("b200mm_preprocess_u8_packed_pil_u8", [c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_int, c_int, c_int, c_int, c_ptr, c_ptr])
declare("b200mm_augment_pil", [c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_int, c_int, c_int, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr])
declare("b200mm_u8_normalize_nchw", [c_ptr, c_ptr, c_int, c_int, c_int, c_ptr, c_ptr, c_ptr, c_ptr])
declare("b200mm_jpeg_parse", [c_ptr, c_longlong, c_ptr])
declare("b200mm_jpeg_entropy_decode", [c_ptr, c_longlong, c_ptr, c_ptr, c_ptr])
declare("b200mm_jpeg_reconstruct", [c_ptr, c_ptr, c_ptr, c_int, c_int, c_int, c_int, c_ptr, c_ptr, c_ptr])
declare("b200mm_jpeg_entropy_decode_sparse", [c_ptr, c_longlong, c_ptr, c_ptr, c_ptr, c_ptr, c_longlong, c_ptr, c_ptr, c_ptr])
declare("b200mm_jpeg_reconstruct_sparse", [c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_int, c_int, c_int, c_int, c_ptr, c_ptr, c_ptr])
declare("b200mm_augment_jitter_rotate", [c_ptr, c_ptr, c_ptr, c_int, c_int, c_int, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr])
declare("b200mm_attention_fwd", [c_ptr, c_ptr, c_ptr, c_ptr, c_int, c_int, c_int, c_float, c_ulonglong, c_ptr])
declare("b200mm_attention_bwd", [c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_int, c_int, c_int, c_float,
                                 c_ulonglong, c_ptr])
declare("b200mm_attention_fwd_mask", [c_ptr, c_ptr, c_ptr, c_ptr, c_int, c_int, c_int, c_float, c_ulonglong, c_ptr, c_ptr])
declare("b200mm_attention_bwd_mask", [c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_int, c_int, c_int, c_float,
                                      c_ulonglong, c_ptr, c_ptr])
declare("b200mm_layernorm_fwd", [c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_int, c_int, c_float, c_float,
                                 c_ulonglong, c_ptr])
declare("b200mm_embed_layernorm_fwd", [c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_int, c_int, c_ptr, c_ptr, c_ptr, c_ptr,
                                       c_ptr, c_ptr, c_int, c_int, c_float, c_float, c_ulonglong, c_ptr])
declare("b200mm_layernorm_bwd", [c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_int, c_int,
                                 c_float, c_ulonglong, c_float, c_ulonglong, c_ptr])
declare("b200mm_embedding_bwd", [c_ptr, c_ptr, c_ptr, c_longlong, c_int, c_int, c_longlong, c_ptr, c_ptr, c_int,
                                 c_int, c_ptr])
declare("b200mm_position_ids", [c_ptr, c_longlong, c_int, c_int, c_ptr, c_ptr])
declare("b200mm_vit_assemble_fwd", [c_ptr, c_ptr, c_ptr, c_ptr, c_int, c_int, c_int, c_ptr])
declare("b200mm_vit_assemble_bwd", [c_ptr, c_ptr, c_ptr, c_ptr, c_int, c_int, c_int, c_ptr])
declare("b200mm_mask_to_bias", [c_ptr, c_ptr, c_longlong, c_ptr])
declare("b200mm_colsum_bf16", [c_ptr, c_longlong, c_int, c_int, c_ptr, c_ptr])
declare("b200mm_head_loss", [c_ptr, c_ptr, c_ptr, c_ptr, c_int, c_int, c_int, c_int, c_float, c_float, c_int,
                             c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr])
declare("b200mm_sumsq_f32", [c_ptr, c_longlong, c_ptr, c_ptr])
declare("b200mm_adam_step", [c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_longlong, c_float, c_float, c_float, c_float,
                             c_float, c_int, c_ptr, c_float, c_float, c_ptr])
declare("b200mm_cast_f32_to_bf16", [c_ptr, c_ptr, c_longlong, c_ptr])
declare("b200mm_scale_cast_f32_to_bf16", [c_ptr, c_ptr, c_longlong, c_float, c_ptr])
declare("b200mm_sumsq_bf16", [c_ptr, c_longlong, c_ptr, c_ptr])
declare("b200mm_adam_step_g16", [c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_longlong, c_float, c_float, c_float, c_float,
                                 c_float, c_int, c_ptr, c_float, c_float, c_ptr])
declare("b200mm_dwconv7x7_nhwc", [c_ptr, c_ptr, c_ptr, c_ptr, c_int, c_int, c_int, c_int, c_ptr])
declare("b200mm_tanh_f32", [c_ptr, c_longlong, c_ptr])
declare("b200mm_set_step_salt_ptr", [c_ptr])
declare("b200mm_adam_step_dyn", [c_ptr, c_ptr, c_int, c_ptr, c_ptr, c_ptr, c_longlong, c_ptr, c_float, c_float, c_float,
                                 c_float, c_ptr, c_ptr, c_float, c_float, c_ptr])
declare("b200mm_step_advance", [c_ptr, c_ptr, c_ptr])
declare("b200mm_gather_rows", [c_ptr, c_ptr, c_int, c_int, c_longlong, c_longlong, c_float, c_ulonglong, c_ptr])
declare("b200mm_scatter_rows", [c_ptr, c_ptr, c_longlong, c_int, c_longlong, c_longlong, c_float, c_ulonglong,
                                c_ptr])
declare("b200mm_batchnorm_fwd", [c_ptr, c_ptr, c_longlong, c_int, c_ptr, c_ptr, c_float, c_float, c_int, c_ptr,
                                 c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr])
declare("b200mm_batchnorm_fwd_stats", [c_ptr, c_ptr, c_longlong, c_int, c_ptr, c_ptr, c_ptr, c_float, c_float, c_int,
                                       c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr])
declare("b200mm_batchnorm_eval", [c_ptr, c_ptr, c_longlong, c_int, c_ptr, c_ptr, c_ptr, c_ptr, c_float, c_int,
                                  c_ptr, c_ptr])
declare("b200mm_batchnorm_bwd", [c_ptr, c_ptr, c_ptr, c_longlong, c_int, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_int,
                                 c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr])
declare("b200mm_maxpool3x3s2_fwd", [c_ptr, c_int, c_int, c_int, c_int, c_ptr, c_ptr, c_ptr])
declare("b200mm_maxpool3x3s2_bwd", [c_ptr, c_ptr, c_int, c_int, c_int, c_int, c_ptr, c_ptr])
declare("b200mm_bn_relu_maxpool_fwd", [c_ptr, c_int, c_int, c_int, c_int, c_ptr, c_ptr, c_ptr, c_float, c_float, c_ptr,
                                       c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr])
declare("b200mm_avgpool_fwd", [c_ptr, c_int, c_int, c_int, c_ptr, c_ptr])
declare("b200mm_avgpool_bwd", [c_ptr, c_int, c_int, c_int, c_ptr, c_ptr])
declare("b200mm_im2col_nhwc", [c_ptr, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_ptr, c_ptr])
declare("b200mm_col2im_nhwc", [c_ptr, c_ptr, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_ptr, c_ptr])
declare("b200mm_im2col_nchw_f32", [c_ptr, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_ptr,
                                   c_ptr])
declare("b200mm_stem_conv_fwd", [c_ptr, c_int, c_int, c_int, c_int, c_ptr, c_int, c_int, c_ptr, c_ptr, c_ptr])
declare("b200mm_stem_conv_wgrad", [c_ptr, c_int, c_int, c_int, c_int, c_ptr, c_int, c_int, c_ptr, c_ptr])
declare("b200mm_subsample_nhwc", [c_ptr, c_int, c_int, c_int, c_int, c_int, c_ptr, c_ptr])
declare("b200mm_upsample_add_nhwc", [c_ptr, c_ptr, c_int, c_int, c_int, c_int, c_int, c_ptr, c_ptr])
declare("b200mm_bn1d_fwd", [c_ptr, c_int, c_longlong, c_int, c_int, c_ptr, c_ptr, c_float, c_float, c_int, c_int, c_ptr,
                            c_longlong, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr])
declare("b200mm_bn1d_bwd", [c_ptr, c_longlong, c_ptr, c_longlong, c_ptr, c_int, c_longlong, c_int, c_int, c_ptr, c_ptr, c_ptr,
                            c_int, c_ptr, c_longlong, c_ptr, c_ptr, c_ptr])
declare("b200mm_softmax_gate_fwd", [c_ptr, c_ptr, c_int, c_int, c_ptr, c_ptr, c_ptr])
declare("b200mm_softmax_gate_bwd", [c_ptr, c_ptr, c_ptr, c_int, c_int, c_ptr, c_ptr, c_ptr])
declare("b200mm_relu_bwd", [c_ptr, c_ptr, c_longlong, c_ptr, c_ptr])
declare("b200mm_head_bn_focal", [c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_int, c_int, c_float,
                                 c_float, c_float, c_float, c_int, c_int, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr,
                                 c_ptr, c_ptr, c_ptr, c_ptr])
