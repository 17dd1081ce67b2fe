"""ctypes binding of csrc/libunetr_b200.so (C ABI declared in include/unetr_b200.h).

There is no CPU implementation behind this module: if the shared library is missing, or the
current device is not an sm_100 GPU, calls raise -- they never fall back.
"""
import ctypes
import os
from ctypes import POINTER, c_char_p, c_float, c_int, c_int32, c_int64, c_size_t, c_uint8, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libunetr_b200.so")
PARAM_COUNT = 164

FLAG_NEED_ENCODER_GRAD = 1
FLAG_HAS_DLOGITS = 2
FLAG_HAS_DENC4 = 4
FLAG_NO_BACKWARD = 16
FLAG_INPLACE_WGRADS = 32
FLAG_WEIGHTS_PACKED = 64


class UnetrConfig(ctypes.Structure):
    _fields_ = [(n, c_int32) for n in (
        "batch", "in_channels", "out_channels", "img0", "img1", "img2", "feature_size", "hidden_size", "mlp_dim",
        "num_heads", "conv_patch_embed", "mode")]


class RankGeom(ctypes.Structure):
    _fields_ = [("src", c_void_p * 4), ("grad", c_void_p * 4),
                ("stride_c", c_int64), ("stride_slice", c_int64), ("stride_f0", c_int64), ("stride_f1", c_int64),
                ("channels", c_int32), ("f0", c_int32), ("f1", c_int32), ("idx", c_int32 * 4), ("temperature", c_float)]


class AugMap(ctypes.Structure):
    _fields_ = [("perm", c_int32 * 3), ("flip", c_int32 * 3), ("shift", c_float)]


class SwGeom(ctypes.Structure):
    _fields_ = [(n, c_int32) for n in (
        "channels", "d", "h", "w", "pad_d", "pad_h", "pad_w", "padded_d", "padded_h", "padded_w", "roi0", "roi1", "roi2")]


# name -> (restype, argtypes); must list every symbol of include/unetr_b200.h (tests/test_capi.py checks)
SIGNATURES = {
    "b200_last_error": (c_char_p, []),
    "b200_device_check": (c_int, []),
    "b200_unetr_create": (c_void_p, [POINTER(UnetrConfig)]),
    "b200_unetr_destroy": (None, [c_void_p]),
    "b200_unetr_workspace_bytes": (c_size_t, [c_void_p, c_int]),
    "b200_unetr_forward": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p]),
    "b200_unetr_backward": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p]),
    "b200_dicece_scratch_bytes": (c_size_t, [c_int, c_int]),
    "b200_dicece_forward": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int64, c_void_p, c_void_p, c_void_p]),
    "b200_dicece_backward": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int64, c_void_p, c_void_p, c_void_p, c_void_p]),
    "b200_ranking_scratch_bytes": (c_size_t, [c_int]),
    "b200_ranking_forward": (c_int, [POINTER(RankGeom), c_void_p, c_void_p, c_void_p]),
    "b200_ranking_backward": (c_int, [POINTER(RankGeom), c_void_p, c_void_p, c_void_p]),
    "b200_sw_gather": (c_int, [c_void_p, c_void_p, POINTER(SwGeom), POINTER(c_int32), c_int, c_float, c_void_p]),
    "b200_sw_accumulate": (c_int, [c_void_p, c_void_p, POINTER(SwGeom), POINTER(c_int32), c_void_p]),
    "b200_sw_accumulate_n": (c_int, [c_void_p, c_void_p, POINTER(SwGeom), POINTER(c_int32), c_int, c_void_p]),
    "b200_sw_finalize": (c_int, [c_void_p, c_void_p, c_void_p, POINTER(SwGeom), c_int, POINTER(c_int32), c_int,
                                 POINTER(c_int32), c_int, POINTER(c_int32), c_int, c_void_p]),
    "b200_sw_finalize_metric": (c_int, [c_void_p, c_void_p, c_void_p, POINTER(SwGeom), c_int, POINTER(c_int32), c_int,
                                        POINTER(c_int32), c_int, POINTER(c_int32), c_int, c_void_p, c_void_p, c_void_p]),
    "b200_sw_accumulate_slab": (c_int, [c_void_p, POINTER(SwGeom), POINTER(c_void_p), POINTER(c_int32), c_int, c_int, c_int, c_void_p]),
    "b200_sw_pack_rows": (c_int, [c_void_p, c_void_p, POINTER(SwGeom), c_int, c_int, c_void_p]),
    "b200_sw_finalize_slab": (c_int, [c_void_p, c_void_p, c_void_p, POINTER(SwGeom), c_int, POINTER(c_int32), c_int, POINTER(c_int32), c_int,
                                      POINTER(c_int32), c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "b200_aug_blocks": (c_int, [c_int64]),
    "b200_aug_index": (c_int, [c_void_p, c_int, c_void_p, c_int, c_float, c_int64, c_void_p, c_void_p, c_void_p]),
    "b200_aug_pick_centers": (c_int, [c_void_p, c_int, c_void_p, c_int, c_float, c_int, c_int, c_int, c_void_p, POINTER(c_int64), c_int,
                                      c_int, c_int, c_int, c_void_p, c_void_p]),
    "b200_aug_crop": (c_int, [c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_void_p, POINTER(AugMap), c_int, c_int, c_int, c_int,
                              c_int, c_void_p, c_void_p, c_void_p]),
    "b200_dicece_sigmoid_forward": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int64, c_void_p, c_void_p, c_void_p]),
    "b200_dicece_sigmoid_backward": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int64, c_void_p, c_void_p, c_void_p, c_void_p]),
    "b200_seg_counts_onehot": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int64, c_void_p, c_void_p]),
    "b200_seg_counts_labels": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int64, c_void_p, c_void_p]),
    "b200_seg_metrics": (c_int, [c_void_p, c_int, c_int, c_int64, c_void_p, c_void_p, c_void_p]),
    "b200_metric_reduce": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "b200_confusion_metric": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p]),
    "b200_unetr_peek": (c_int, [c_void_p, c_char_p, c_void_p, c_size_t]),
    "b200_launch_count": (ctypes.c_ulonglong, []),
    "b200_prof_enable": (None, [c_int]),
    "b200_prof_report": (c_int, [ctypes.c_char_p, c_int]),
    "b200_test_tc_conv": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_int, c_int,
                                  c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "b200_test_tc_conv_fused": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_int,
                                        c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "b200_test_tc_wgrad": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                                   c_void_p, c_void_p]),
    "b200_test_layernorm_bwd": (c_int, [c_void_p] * 9 + [c_int, c_int, c_int, c_void_p]),
    "b200_test_instnorm_bwd": (c_int, [c_int] + [c_void_p] * 6 + [c_int, c_int, c_int64, c_void_p, c_void_p, c_void_p, c_int, c_void_p]),
    "b200_test_head_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int64, c_void_p, c_void_p, c_void_p, c_int, c_void_p]),
    "b200_test_tc_conv_dgrad_normbwd": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                                                POINTER(c_int), c_void_p, c_void_p]),
    "b200_test_set_debug_buffer": (None, [c_void_p]),
    "b200_adamw_chunk": (ctypes.c_long, []),
    "b200_adamw_step": (c_int, [c_void_p, c_void_p, c_int, ctypes.c_float, ctypes.c_float, ctypes.c_float, ctypes.c_float, ctypes.c_float,
                                c_int, c_void_p]),
    "b200_adamw_step_capturable": (c_int, [c_void_p, c_void_p, c_int, ctypes.c_float, ctypes.c_float, ctypes.c_float, ctypes.c_float,
                                           ctypes.c_float, c_int, c_void_p, c_void_p]),
    "b200_unetr_packed_bytes": (c_size_t, [c_void_p]),
    "b200_unetr_set_packed_weights": (None, [c_void_p, c_void_p]),
    "b200_unetr_packed_cast_offset": (c_int64, [c_void_p, c_int]),
    "b200_unetr_pack_convs": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p]),
    "b200_unetr_set_grad_events": (None, [c_void_p, ctypes.POINTER(c_void_p), c_int]),
    "b200_trace_begin": (None, [c_void_p, c_int]),
    "b200_trace_count": (c_int, []),
    "b200_trace_tags": (c_int, [ctypes.c_char_p, c_int]),
    "b200_test_tc_attention": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_float, c_void_p]),
    "b200_test_tc_attention_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_float, c_void_p]),
    "b200_test_tc_attention_bwd_kv": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "b200_test_tc_gemm": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "b200_test_tc_gemm_grouped": (c_int, [POINTER(c_void_p), POINTER(c_void_p), POINTER(c_void_p), POINTER(c_int), POINTER(c_int), POINTER(c_int),
                                          c_int, c_int, c_void_p]),
}

_lib = None
_device_ok = set()


def load():
    """Load the shared library (raises with build instructions if it is absent)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python __graft_entry__.py build` (make -C csrc). "
                "This package has no CPU or PyTorch fallback.")
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def last_error() -> str:
    msg = load().b200_last_error()
    return msg.decode() if msg else ""


def check(rc: int, what: str):
    if rc != 0:
        raise RuntimeError(f"{what} failed (code {rc}): {last_error()}")


def require_device(tensor):
    """All entry points take CUDA tensors on an sm_100 device.  One device per process (the multi-GPU design is one process per
    GPU, parallel.py): the library caches per-device launch attributes and the SM count for the first device it sees, and every
    call launches on the current stream of the current device, so a tensor on another device is refused rather than silently
    launched into the wrong context."""
    import torch

    if not tensor.is_cuda:
        raise RuntimeError("b200 UNETR kernels need CUDA tensors on a B200 (sm_100a); there is no CPU path")
    dev = tensor.device.index if tensor.device.index is not None else torch.cuda.current_device()
    if dev not in _device_ok:
        if _device_ok:
            raise RuntimeError(f"libunetr_b200 is bound to cuda:{next(iter(_device_ok))} in this process; got a tensor on cuda:{dev} "
                               "(run one process per GPU: torchrun / parallel.init_from_env)")
        with torch.cuda.device(dev):
            check(load().b200_device_check(), "b200_device_check")
        _device_ok.add(dev)
    if dev != torch.cuda.current_device():
        raise RuntimeError(f"tensor on cuda:{dev} but the current device is cuda:{torch.cuda.current_device()}: call "
                           "torch.cuda.set_device first (kernels launch on the current device's stream)")
    return dev


def stream_ptr():
    import torch

    return c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    return c_void_p(t.data_ptr()) if t is not None else c_void_p(0)


def prof_report() -> dict:
    """{tag: (total_ms, launches)} since the last call (synchronises the device)."""
    buf = ctypes.create_string_buffer(1 << 16)
    load().b200_prof_report(buf, len(buf))
    out = {}
    for line in buf.value.decode().splitlines():
        tag, ms, n = line.rsplit(" ", 2)
        out[tag] = (float(ms), int(n))
    return out
