"""ctypes binding of libmultibox_b200.so (the C ABI in include/multibox_b200.h).

There is no CPU fallback: if the shared library is missing and cannot be built
(no nvcc), importing the CUDA entry points raises.  PyTorch only owns the device
memory and the stream; every pointer handed over is ``tensor.data_ptr()``.
"""
import ctypes
import os

from . import _build

_c_void_p = ctypes.c_void_p
_c_int = ctypes.c_int
_c_uint = ctypes.c_uint
_c_float = ctypes.c_float
_c_size_t = ctypes.c_size_t

# flags / codes (mirror include/multibox_b200.h)
FLAG_LOGITS = 1
FLAG_BOUNDARY = 2
FLAG_GENERIC = 4
FLAG_AR_DEFERRED = 8
FLAG_STATIC = 16
FLAG_HOST_RESULTS = 32
FLAG_PDL = 64
FLAG_WARPS_SHIFT = 8
FLAG_COLS_SHIFT = 16
STATUS_INVALID_COST = 1
STATUS_INFEASIBLE = 2
STATUS_BAD_NUM_GT = 4
STATUS_AR_TIMEOUT = 8
MAX_PEERS = 8
RESULT_WORDS = 16

EXPORTS = ("mbx_version", "mbx_last_error", "mbx_device_info",
           "mbx_match_workspace_bytes", "mbx_match_loss", "mbx_match_loss_ragged", "mbx_match_loss_heads",
           "mbx_allreduce_buffer_bytes", "mbx_allreduce_config", "mbx_match_loss_allreduce", "mbx_allreduce_flush",
           "mbx_match_plan_create", "mbx_match_plan_launch", "mbx_match_plan_launch_staged", "mbx_match_plan_destroy",
           "mbx_detect_workspace_bytes", "mbx_detect", "mbx_detect_heads",
           "mbx_filter_proposals", "mbx_convert_proposals",
           "mbx_debug_nplog", "mbx_debug_cost_matrix", "mbx_debug_sqrt_mismatches", "mbx_debug_fastlog_violations")

_lib = None
MAX_HEADS = 8


class Heads(ctypes.Structure):
    """ctypes image of `mbx_heads` (include/multibox_b200.h)."""
    _fields_ = [("num_heads", ctypes.c_int32),
                ("head_priors", ctypes.c_int32 * MAX_HEADS),
                ("locations", ctypes.c_void_p * MAX_HEADS),
                ("confidences", ctypes.c_void_p * MAX_HEADS),
                ("d_locations", ctypes.c_void_p * MAX_HEADS),
                ("d_confidences", ctypes.c_void_p * MAX_HEADS)]


class MultiboxLibraryError(RuntimeError):
    pass


def library_path():
    return _build.LIB


def load():
    """Loads (building first if the sources are newer and nvcc exists) the library."""
    global _lib
    if _lib is not None:
        return _lib
    path = _build.LIB
    if _build.needs_build():
        try:
            _build.nvcc_path()
            have_nvcc = True
        except RuntimeError:
            have_nvcc = False
        if have_nvcc:
            # sources are newer than the library and a compiler exists: a compile error must surface,
            # never be hidden behind a stale binary
            try:
                _build.build()
            except Exception as e:
                raise MultiboxLibraryError("building libmultibox_b200.so failed (%s); refusing to run a stale or "
                                           "missing library; there is no CPU fallback" % e)
        elif not os.path.isfile(path):
            raise MultiboxLibraryError(
                "libmultibox_b200.so is missing and nvcc is not available; there is no CPU fallback")
        # (no nvcc on this box and a prebuilt .so travelled with the repo: use it)
    lib = ctypes.CDLL(path)
    lib.mbx_version.restype = _c_int
    lib.mbx_last_error.restype = ctypes.c_char_p
    lib.mbx_device_info.restype = _c_int
    lib.mbx_device_info.argtypes = [ctypes.POINTER(_c_int), ctypes.POINTER(_c_int)]
    lib.mbx_match_workspace_bytes.restype = _c_size_t
    lib.mbx_match_workspace_bytes.argtypes = [_c_int, _c_int, _c_int]
    lib.mbx_match_loss.restype = _c_int
    lib.mbx_match_loss.argtypes = [
        _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p,      # locations, confidences, gt, num_gt, priors
        _c_int, _c_int, _c_int, _c_float, _c_uint,                  # B, P, M, alpha, flags
        _c_void_p, _c_void_p, _c_void_p, _c_void_p,                 # mask, matched_gt_idx, stacked_gt, n_stacked
        _c_void_p, _c_void_p, _c_void_p, _c_void_p,                 # d_loc, d_conf, conf_out, results
        _c_void_p, _c_size_t, _c_void_p]                            # workspace, bytes, stream
    lib.mbx_match_loss_ragged.restype = _c_int
    lib.mbx_match_loss_ragged.argtypes = lib.mbx_match_loss.argtypes    # gt_flat, gt_row_offsets in place of gt, num_gt
    lib.mbx_match_loss_heads.restype = _c_int
    lib.mbx_match_loss_heads.argtypes = [
        ctypes.POINTER(Heads), _c_void_p, _c_void_p, _c_void_p, _c_void_p,   # heads, gt, num_gt, gt_row_offsets, priors
        _c_int, _c_int, _c_int, _c_float, _c_uint,                  # B, P, M, alpha, flags
        _c_void_p, _c_void_p, _c_void_p, _c_void_p,                 # mask, matched_gt_idx, stacked_gt, n_stacked
        _c_void_p, _c_void_p,                                       # conf_out, results
        _c_void_p, _c_size_t, _c_void_p]                            # workspace, bytes, stream
    lib.mbx_allreduce_buffer_bytes.restype = _c_size_t
    lib.mbx_allreduce_buffer_bytes.argtypes = []
    lib.mbx_allreduce_config.restype = _c_int
    lib.mbx_allreduce_config.argtypes = [_c_int, _c_int]
    lib.mbx_match_loss_allreduce.restype = _c_int
    lib.mbx_match_loss_allreduce.argtypes = lib.mbx_match_loss.argtypes[:-1] + [_c_void_p, _c_int, _c_int, _c_void_p]
    lib.mbx_match_plan_create.restype = _c_int
    lib.mbx_match_plan_create.argtypes = [ctypes.POINTER(_c_void_p)] + lib.mbx_match_loss.argtypes[:-1] + \
        [_c_void_p, _c_int, _c_int]
    lib.mbx_match_plan_launch.restype = _c_int
    lib.mbx_match_plan_launch.argtypes = [_c_void_p, _c_void_p]
    lib.mbx_match_plan_launch_staged.restype = _c_int
    lib.mbx_match_plan_launch_staged.argtypes = [_c_void_p, _c_void_p, _c_void_p, _c_size_t, _c_void_p]
    lib.mbx_match_plan_destroy.restype = None
    lib.mbx_match_plan_destroy.argtypes = [_c_void_p]
    lib.mbx_allreduce_flush.restype = _c_int
    lib.mbx_allreduce_flush.argtypes = [_c_void_p, _c_void_p, _c_size_t, _c_void_p, _c_int, _c_int, _c_void_p]
    lib.mbx_detect_workspace_bytes.restype = _c_size_t
    lib.mbx_detect_workspace_bytes.argtypes = [_c_int, _c_int, _c_int]
    lib.mbx_detect.restype = _c_int
    lib.mbx_detect.argtypes = [
        _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p,      # locations, confidences, priors, restrictions, max_to_keep
        _c_void_p, _c_void_p, _c_void_p, _c_void_p,                 # offsets, patch_dims, image_dims, is_flipped
        _c_int, _c_int, _c_int, _c_float, _c_uint,                  # B, P, k_max, nms_iou, flags
        _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p,      # out_boxes, out_patch_boxes, out_scores, out_idx, out_count
        _c_void_p, _c_size_t, _c_void_p]                            # workspace, bytes, stream
    lib.mbx_detect_heads.restype = _c_int
    lib.mbx_detect_heads.argtypes = [ctypes.POINTER(Heads)] + lib.mbx_detect.argtypes[2:]
    lib.mbx_filter_proposals.restype = _c_int
    lib.mbx_filter_proposals.argtypes = [
        _c_void_p, _c_void_p, _c_void_p, _c_int, _c_int,            # bboxes, confidences, restrictions, B, P
        _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p]      # out_bboxes, out_conf, out_idx, out_count, stream
    lib.mbx_convert_proposals.restype = _c_int
    lib.mbx_convert_proposals.argtypes = [
        _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p,      # bboxes, offsets, patch_dims, image_dims, is_flipped
        _c_void_p, _c_int, _c_int, _c_void_p, _c_void_p]            # counts, B, K, out_boxes, stream
    lib.mbx_debug_nplog.restype = _c_int
    lib.mbx_debug_nplog.argtypes = [_c_void_p, _c_void_p, ctypes.c_longlong, _c_void_p]
    lib.mbx_debug_sqrt_mismatches.restype = _c_int
    lib.mbx_debug_sqrt_mismatches.argtypes = [_c_uint, _c_uint, _c_void_p, _c_void_p]
    lib.mbx_debug_fastlog_violations.restype = _c_int
    lib.mbx_debug_fastlog_violations.argtypes = [_c_uint, _c_uint, _c_void_p, _c_void_p]
    lib.mbx_debug_cost_matrix.restype = _c_int
    lib.mbx_debug_cost_matrix.argtypes = [_c_void_p, _c_void_p, _c_void_p, _c_int, _c_int, _c_float,
                                          _c_void_p, _c_void_p]
    _lib = lib
    return lib


def last_error():
    return load().mbx_last_error().decode("utf-8", "replace")


def check(rc, what):
    if rc != 0:
        raise MultiboxLibraryError("%s failed (code %d): %s" % (what, rc, last_error()))


def ptr(t):
    """data_ptr of a tensor or None."""
    return None if t is None else t.data_ptr()
