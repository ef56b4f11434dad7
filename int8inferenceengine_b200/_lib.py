"""ctypes binding of libi8ie_sm100.so (include/i8ie_sm100.h). There is NO CPU fallback:
if the library is missing or a call fails, this raises."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libi8ie_sm100.so")

# every symbol include/i8ie_sm100.h declares
SYMBOLS = [
    "i8ie_last_error", "i8ie_version", "i8ie_device_check", "i8ie_launch_count",
    "i8ie_quantize_f32_u8", "i8ie_quantize_nchw_f32_nhwc_u8", "i8ie_dequantize_u8_f32",
    "i8ie_dequantize_rows_u8_f32", "i8ie_downscale_s32_u8", "i8ie_minmax_workspace_bytes",
    "i8ie_minmax_f32", "i8ie_range_from_minmax_host", "i8ie_relu_u8", "i8ie_maxpool_u8_nhwc",
    "i8ie_u8_nchw_to_nhwc", "i8ie_u8_nhwc_to_nchw", "i8ie_quantize_weight_host", "i8ie_zp_offsets",
    "i8ie_pack_conv_weight", "i8ie_conv2d_plan_create", "i8ie_conv2d_plan_destroy",
    "i8ie_conv2d_plan_impl", "i8ie_conv2d_u8", "i8ie_conv2d_f32_u8", "i8ie_fc_u8", "i8ie_debug_tc_error",
    "i8ie_quantize_f32_u8_indirect", "i8ie_quantize_nchw_f32_nhwc_u8_indirect", "i8ie_copy_indirect",
    "i8ie_conv2d_f32_u8_indirect", "i8ie_top1_chunk_bytes", "i8ie_top1_pack", "i8ie_top1_unpack",
    "i8ie_conv2d_f32", "i8ie_linear_f32", "i8ie_relu_f32", "i8ie_maxpool_f32_nchw",
    "i8ie_tc_error_poll", "i8ie_peer_exchange_bytes", "i8ie_peer_alloc", "i8ie_peer_open", "i8ie_peer_close",
    "i8ie_peer_free", "i8ie_top1_pack_push", "i8ie_top1_wait_unpack",
    "i8ie_maxpool_u8_nhwc_padded", "i8ie_conv2d_row_mode_cp", "i8ie_conv2d_plan_set_channel_scales", "i8ie_fc_u8_pc",
    "i8ie_fc_u8_deq", "i8ie_fc_weight_tiled_bytes", "i8ie_fc_weight_tiled_attach", "i8ie_fc_weight_tiled_detach",
]

_lib = None


class I8ieError(RuntimeError):
    pass


def load():
    """Load the CUDA library (building it in-tree first if nvcc is available and it is stale)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        try:
            from . import build as _build
            _build.build()
        except Exception as e:  # noqa: BLE001
            raise I8ieError(
                f"libi8ie_sm100.so is missing and could not be built ({e}); the i8ie B200 backend has "
                "no CPU fallback — run `python -m int8inferenceengine_b200.build`") from e
    L = C.CDLL(LIB_PATH)
    i, i64, f, vp = C.c_int, C.c_int64, C.c_float, C.c_void_p
    L.i8ie_last_error.restype = C.c_char_p
    L.i8ie_version.restype = C.c_char_p
    L.i8ie_launch_count.restype = i64
    L.i8ie_minmax_workspace_bytes.restype = i64
    L.i8ie_quantize_f32_u8.argtypes = [vp, vp, i64, f, i, vp]
    L.i8ie_quantize_nchw_f32_nhwc_u8.argtypes = [vp, vp, i, i, i, i, i, f, i, vp]
    L.i8ie_dequantize_u8_f32.argtypes = [vp, vp, i64, f, i, vp]
    L.i8ie_dequantize_rows_u8_f32.argtypes = [vp, vp, i, i, i, f, i, vp]
    L.i8ie_downscale_s32_u8.argtypes = [vp, vp, i64, f, f, f, i, vp]
    L.i8ie_minmax_f32.argtypes = [vp, i64, vp, vp, vp]
    L.i8ie_range_from_minmax_host.argtypes = [f, f, C.POINTER(f), C.POINTER(C.c_uint8)]
    L.i8ie_relu_u8.argtypes = [vp, vp, i64, i, vp]
    L.i8ie_maxpool_u8_nhwc.argtypes = [vp, vp, i, i, i, i, i, i, i, i, vp]
    L.i8ie_maxpool_u8_nhwc_padded.argtypes = [vp, vp, i, i, i, i, i, i, i, i, i, i, vp]
    L.i8ie_conv2d_row_mode_cp.argtypes = [i, i, i, i, i, i, i]
    L.i8ie_u8_nchw_to_nhwc.argtypes = [vp, vp, i, i, i, i, i, i, vp]
    L.i8ie_u8_nhwc_to_nchw.argtypes = [vp, vp, i, i, i, i, i, vp]
    L.i8ie_quantize_weight_host.argtypes = [vp, i64, vp, i64, vp, vp, C.POINTER(f)]
    L.i8ie_zp_offsets.argtypes = [vp, vp, i, i, i, f, i, vp, vp, vp]
    L.i8ie_pack_conv_weight.argtypes = [vp, vp, i, i, i, i, i, i, vp]
    L.i8ie_conv2d_plan_create.argtypes = [i, i, i, i, i, i, i, i, i, i, i, vp, i, i]
    L.i8ie_conv2d_plan_create.restype = vp
    L.i8ie_conv2d_plan_destroy.argtypes = [vp]
    L.i8ie_conv2d_plan_destroy.restype = None
    L.i8ie_conv2d_plan_impl.argtypes = [vp]
    L.i8ie_conv2d_u8.argtypes = [vp, vp, vp, vp, f, f, f, i, i, i, vp, vp]
    L.i8ie_fc_u8.argtypes = [vp, i, vp, i, i, vp, i, i, i, i, vp, vp, f, f, f, i, i, vp, i, vp]
    L.i8ie_conv2d_plan_set_channel_scales.argtypes = [vp, vp, f, f]
    L.i8ie_fc_u8_pc.argtypes = [vp, i, vp, i, i, vp, i, i, i, i, vp, vp, f, vp, f, f, f, i, i, vp, i, vp]
    L.i8ie_fc_u8_deq.argtypes = [vp, i, vp, i, i, vp, i, i, i, i, vp, vp, f, f, vp, f, f, f, i, i, i, vp, vp]
    L.i8ie_fc_weight_tiled_bytes.argtypes = [i, i]
    L.i8ie_fc_weight_tiled_bytes.restype = i64
    L.i8ie_fc_weight_tiled_attach.argtypes = [vp, i, i, vp, vp]
    L.i8ie_fc_weight_tiled_detach.argtypes = [vp]
    L.i8ie_conv2d_f32_u8.argtypes = [vp, vp, f, i, vp, vp, f, f, i, i, vp, vp]
    L.i8ie_conv2d_f32_u8_indirect.argtypes = [vp, vp, f, i, vp, vp, f, f, i, i, vp, vp]
    L.i8ie_quantize_f32_u8_indirect.argtypes = [vp, vp, i64, f, i, vp]
    L.i8ie_quantize_nchw_f32_nhwc_u8_indirect.argtypes = [vp, vp, i, i, i, i, i, f, i, vp]
    L.i8ie_copy_indirect.argtypes = [vp, vp, i64, vp]
    L.i8ie_top1_chunk_bytes.argtypes = [i, i]
    L.i8ie_top1_chunk_bytes.restype = i64
    L.i8ie_top1_pack.argtypes = [vp, vp, i, i, vp, vp]
    L.i8ie_top1_unpack.argtypes = [vp, i, i64, vp, vp, vp]
    L.i8ie_debug_tc_error.argtypes = [i]
    L.i8ie_tc_error_poll.argtypes = [i]
    L.i8ie_peer_exchange_bytes.argtypes = [i, i64]
    L.i8ie_peer_exchange_bytes.restype = i64
    L.i8ie_peer_alloc.argtypes = [i64, C.POINTER(vp), vp]
    L.i8ie_peer_open.argtypes = [vp, C.POINTER(vp)]
    L.i8ie_peer_close.argtypes = [vp]
    L.i8ie_peer_free.argtypes = [vp]
    L.i8ie_top1_pack_push.argtypes = [vp, vp, i, i, C.POINTER(vp), i, i, i64, vp, vp]
    L.i8ie_top1_wait_unpack.argtypes = [vp, i, i64, vp, vp, vp, vp]
    L.i8ie_conv2d_f32.argtypes = [vp, vp, vp, vp, i, i, i, i, i, i, i, i, i, vp, vp]
    L.i8ie_linear_f32.argtypes = [vp, vp, vp, vp, i, i, i, vp, vp]
    L.i8ie_relu_f32.argtypes = [vp, vp, i64, vp]
    L.i8ie_maxpool_f32_nchw.argtypes = [vp, vp, i, i, i, i, i, i, vp]
    _lib = L
    return L


def last_error():
    return load().i8ie_last_error().decode(errors="replace")


def check(rc, what=""):
    if rc != 0:
        raise I8ieError(f"{what or 'i8ie call'} failed (code {rc}): {last_error()}")


_TC_ROLES = {1: "TMA producer waiting for a free operand stage", 2: "MMA issuer waiting for operands",
             3: "epilogue waiting for an accumulator", 4: "MMA issuer waiting for a drained accumulator",
             5: "stem MMA issuer waiting for the weights", 6: "stem converter waiting for fp32 rows",
             7: "result exchange waiting for a peer rank's chunk",
             8: "split-K fc CTA waiting for its sibling CTAs' partial tiles"}


def check_tc_error():
    """Raises if a tensor-core kernel recorded a pipeline fault (mbarrier wait timeout) on the
    current device. Reads a host-mapped flag (no CUDA call); call after a synchronisation."""
    if _lib is None:
        return
    code = int(_lib.i8ie_tc_error_poll(1))
    if code != 0:
        raise I8ieError(f"a kernel hit a bounded-wait timeout (role {code}: "
                        f"{_TC_ROLES.get(code, 'unknown')}); its output is invalid")


def launch_count():
    return int(load().i8ie_launch_count())
