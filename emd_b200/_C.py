"""ctypes binding of ``libemd_b200.so`` (the C ABI declared in ``include/emd_b200.h``).

Every entry point takes raw device pointers, sizes and a ``cudaStream_t``; all
memory is allocated and owned by PyTorch on the Python side.  There is no CPU
fallback: a missing library, a non-CUDA tensor or a non-zero status raises.
"""
from __future__ import annotations

import ctypes
from ctypes import c_float, c_int, c_int64, c_size_t, c_void_p
from typing import Optional

import torch

from . import build as _build

_lib = None

P = c_void_p
_SIGS = {
    # name: (restype, argtypes)
    "emd_last_error_string": (ctypes.c_char_p, []),
    "emd_abi_version": (c_int, []),
    "emd_device_check": (c_int, []),
    "emd_launch_count": (ctypes.c_longlong, []),
    "emd_kernel_id_count": (c_int, []),
    "emd_kernel_name": (ctypes.c_char_p, [c_int]),
    "emd_profile_enable": (None, [c_int]),
    "emd_profile_collect": (c_int, [ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_longlong)]),
    "emd_raster_set_counters": (None, [P]),
    "emd_fp32_probe": (c_int, [c_int, c_int, P, P]),
    "emd_fp32_probe_lane_instructions": (c_int64, [c_int, c_int]),
    "emd_projection_fwd": (c_int, [P, P, P, P, P, c_int64, c_int64, c_int, c_int, c_float, c_float, c_float, c_float,
                                   c_int, c_int, P, P, P, P, P, P, P]),
    "emd_projection_bwd": (c_int, [P, P, P, P, P, c_int64, c_int64, c_int, c_int, c_float, c_float, c_float, c_float,
                                   P, P, P, P, P, P, P, P]),
    "emd_scan_workspace_bytes": (c_size_t, [c_int64]),
    "emd_cumsum_i32_i64": (c_int, [P, P, c_int64, P, P, c_size_t, P]),
    "emd_exclusive_scan_u32": (c_int, [P, P, c_int64, P, c_size_t, P]),
    "emd_exclusive_scan_u8_u32": (c_int, [P, P, c_int64, P, P, c_size_t, P]),
    "emd_isect_emit": (c_int, [P, P, P, P, c_int64, c_int64, c_int, c_int, c_int, P, P, P]),
    "emd_isect_offsets": (c_int, [P, c_int64, c_int64, c_int, c_int, c_int, P, P]),
    "emd_radix_sort_workspace_bytes": (c_size_t, [c_int64]),
    "emd_radix_sort_pairs": (c_int, [P, P, P, P, c_int64, c_int, c_int, P, c_size_t, ctypes.POINTER(c_int), P]),
    "emd_raster_pack": (c_int, [P, P, P, c_int, P, c_int, c_int, P, c_int, P, c_int64, c_int64, P, P]),
    "emd_linear_bwd_workspace_bytes": (c_size_t, [c_int64, c_int, c_int]),
    "emd_linear_fwd": (c_int, [P, P, P, c_int64, c_int, c_int, c_int, c_int, P, P]),
    "emd_linear_fwd_tc": (c_int, [P, P, P, c_int64, c_int, c_int, c_int, c_int, P, P]),
    "emd_linear_bwd": (c_int, [P, P, P, P, c_int64, c_int, c_int, c_int, c_int, P, P, P, P, c_size_t, P]),
    "emd_linear_bwd_tc": (c_int, [P, P, P, P, c_int64, c_int, c_int, c_int, c_int, P, P, P, P, c_size_t, P]),
    "emd_temb_fwd": (c_int, [P, c_int, c_int, P, c_int, P, P]),
    "emd_temb_bwd": (c_int, [P, c_int, c_int, P, c_int, P, P, P, P]),
    "emd_hexplane_fwd": (c_int, [P, ctypes.POINTER(c_int64), ctypes.POINTER(c_int), c_int, c_int, ctypes.POINTER(c_float),
                                 P, P, c_int, c_int64, P, P]),
    "emd_hexplane_fwd_ld": (c_int, [P, ctypes.POINTER(c_int64), ctypes.POINTER(c_int), c_int, c_int, ctypes.POINTER(c_float),
                                    P, P, c_int, c_int64, P, c_int64, P]),
    "emd_hexplane_bwd_ld": (c_int, [P, ctypes.POINTER(c_int64), ctypes.POINTER(c_int), c_int, c_int, ctypes.POINTER(c_float),
                                    P, P, c_int, c_int64, P, c_int64, P, P, P, P, c_size_t, P]),
    "emd_s3g_apply_blocks": (c_int64, [c_int64]),
    "emd_s3g_apply_fwd": (c_int, [P] * 10 + [c_int64] + [P] * 4 + [P]),
    "emd_s3g_apply_bwd": (c_int, [P] * 10 + [c_int64] + [P] * 8 + [P]),
    "emd_hexplane_bwd_workspace_bytes": (c_size_t, [c_int64]),
    "emd_hexplane_bwd": (c_int, [P, ctypes.POINTER(c_int64), ctypes.POINTER(c_int), c_int, c_int, ctypes.POINTER(c_float),
                                 P, P, c_int, c_int64, P, P, P, P, P, c_size_t, P]),
    "emd_voxel_lbs_fwd": (c_int, [P, P, P, P, c_float, c_int, c_int, c_int, c_int, c_int, c_int, P, c_int64, P, P]),
    "emd_voxel_lbs_bwd": (c_int, [P, P, P, P, c_float, c_int, c_int, c_int, c_int, c_int, c_int, P, c_int64, P, P, P, P]),
    "emd_dense_fwd": (c_int, [P, c_int64, P, P, c_int64, c_int, c_int, c_int, P, c_int64, P]),
    "emd_dense_bwd_workspace_bytes": (c_size_t, [c_int64, c_int, c_int]),
    "emd_dense_bwd": (c_int, [P, c_int64, P, P, c_int64, c_int64, c_int, c_int, P, c_int64, c_int, c_int, P, c_int64, P, P,
                              P, c_size_t, P]),
    "emd_dense_tc_enabled": (c_int, []),
    "emd_dense_set_tc": (None, [c_int]),
    "emd_deform_input_fwd": (c_int, [P, P, P, P, c_float, c_int, c_int, c_int, c_int64, P, c_int64, P, c_int64, P]),
    "emd_deform_embed_grad_workspace_bytes": (c_size_t, [c_int, c_int, c_int64]),
    "emd_deform_embed_grad": (c_int, [P, P, c_int, P, P, c_int, c_int64, P, c_size_t, P, P]),
    "emd_deform_apply_fwd": (c_int, [P, P, P, c_int, c_int64, P, P, P]),
    "emd_deform_apply_bwd": (c_int, [P, P, P, c_int, c_int64, P, P, P, P]),
    "emd_adam_max_tensors": (c_int, []),
    "emd_adam_step": (c_int, [ctypes.POINTER(P)] * 4 + [ctypes.POINTER(c_int64)] + [ctypes.POINTER(ctypes.c_double)] * 5
                      + [ctypes.POINTER(c_int64), c_int, ctypes.c_double, P]),
    "emd_densify_stats": (c_int, [P, P, c_int64, c_int64, c_float, c_float, c_float, c_int, P, P, P, P]),
    "emd_image_loss_partials_floats": (c_int64, [c_int, c_int, c_int]),
    "emd_image_loss_fwd": (c_int, [P] * 8 + [c_int, c_int, c_int, P, ctypes.POINTER(c_float)] + [P] * 4 + [P]),
    "emd_image_loss_bwd": (c_int, [P] * 8 + [c_int, c_int, c_int, P, ctypes.POINTER(c_float)] + [P] * 7 + [P]),
    "emd_tile_order": (c_int, [P, c_int64, c_int64, P, P, P, P, P]),
    "emd_raster_max_ctas": (c_int64, [c_int64, c_int64]),
    "emd_raster_segment_size": (c_int, []),
    "emd_raster_checkpoint_floats": (c_int, []),
    "emd_raster_segout_floats": (c_int, []),
    "emd_raster_segment_slots": (c_int64, [c_int64]),
    "emd_raster_sort_records": (c_int, [P, P, P, P, P, c_int64, c_int, c_int, c_int, c_int, P, P, P]),
    "emd_rasterize_fwd": (c_int, [P, P, P, P, P, P, c_int64, c_int64, c_int, c_int, c_int, c_int, c_int, c_int, c_int, P, P, P,
                                  P, P, P, P]),
    "emd_dg_preprocess_fwd": (c_int, [P] * 4 + [ctypes.POINTER(c_float)] * 3 + [c_float, c_float, c_int, c_int, c_float,
                                                                              c_int, c_int, c_int64] + [P] * 7 + [P]),
    "emd_dg_preprocess_bwd": (c_int, [P] * 4 + [ctypes.POINTER(c_float)] * 3 + [c_float, c_float, c_int, c_int, c_float,
                                                                              c_int, c_int, c_int64] + [P] * 10 + [P]),
    "emd_dg_isect_emit": (c_int, [P, P, P, P, c_int64, c_int, c_int, P, P, P]),
    "emd_rasterize_bwd_workspace_bytes": (c_size_t, [c_int64]),
    "emd_sh_fwd": (c_int, [c_int, P, P, c_int64, c_int, P, P]),
    "emd_sh_bwd": (c_int, [c_int, P, c_int64, c_int, P, P, P]),
    "emd_activate_fwd": (c_int, [P] * 8 + [ctypes.POINTER(c_float), c_int, c_int64, c_int, c_int] + [P] * 5 + [P]),
    "emd_activate_bwd": (c_int, [P] * 8 + [ctypes.POINTER(c_float), c_int, c_int64, c_int, c_int] + [P] * 11 + [P]),
    "emd_rigid_chunk_size": (c_int, []),
    "emd_rigid_param_count": (c_int, [c_int, c_int]),
    "emd_rigid_deform_fwd": (c_int, [P] * 11 + [c_int64, c_int, c_int, c_int, c_int, c_float, c_int, c_int, c_int]
                             + [P] * 5 + [P]),
    "emd_rigid_deform_bwd": (c_int, [P] * 10 + [c_int64, c_int, c_int, c_int, c_int, c_float, c_int, c_int, c_int]
                             + [P] * 15 + [P]),
    "emd_smpl_param_count": (c_int, [c_int, c_int]),
    "emd_smpl_max_chunks": (c_int, [c_int]),
    "emd_smpl_reduce_width": (c_int, []),
    "emd_smpl_deform_fwd": (c_int, [P] * 12 + [c_int] * 5 + [c_float, c_int, c_int] + [P] * 5 + [P]),
    "emd_smpl_deform_bwd": (c_int, [P] * 11 + [c_int] * 5 + [c_float, c_int, c_int] + [P] * 14 + [P]),
    "emd_smpl_weight_grad": (c_int, [P, P, P, P, P, c_int, c_int, P, P, P, P]),
    "emd_rasterize_bwd": (c_int, [P, P, P, P, P, c_int64, c_int64, c_int64, c_int64, c_int, c_int, c_int, c_int, c_int, c_int,
                                  c_int, P, P, P, P, P, P, P, P, P, c_int, c_int, P, P, P, P, P, P, P, c_size_t, P]),
}


class EmdError(RuntimeError):
    pass


def lib() -> ctypes.CDLL:
    """Load (building first if the .so is absent and nvcc is available)."""
    global _lib
    if _lib is not None:
        return _lib
    path = _build.LIB
    if not path.exists():
        try:
            _build.build()
        except Exception as e:  # noqa: BLE001
            raise EmdError(
                f"emd_b200: {path} is missing and could not be built ({e}). There is no CPU fallback; "
                "run `python -m emd_b200.build` with the CUDA 12.9 toolkit."
            ) from e
    try:
        L = ctypes.CDLL(str(path))
    except OSError as e:
        raise EmdError(f"emd_b200: cannot load {path}: {e}. There is no CPU fallback.") from e
    for name, (res, args) in _SIGS.items():
        if not hasattr(L, name):
            raise EmdError(f"emd_b200: {path} does not export {name}; rebuild with `python -m emd_b200.build --force`")
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


def declared_symbols():
    return sorted(_SIGS)


def check(status: int, what: str) -> None:
    if status != 0:
        msg = lib().emd_last_error_string().decode("utf-8", "replace")
        raise EmdError(f"{what} failed with status {status}: {msg}")


def ptr(t: Optional[torch.Tensor], dtype=None, name: str = "tensor") -> Optional[int]:
    """Device pointer of a contiguous CUDA tensor (None passes through as NULL)."""
    if t is None:
        return None
    if t.is_cuda and t.is_contiguous() and (dtype is None or t.dtype == dtype):   # the hot path: one branch
        return t.data_ptr()
    if not t.is_cuda:
        raise EmdError(f"emd_b200: {name} must be a CUDA tensor (got {t.device}); there is no CPU path")
    if not t.is_contiguous():
        raise EmdError(f"emd_b200: {name} must be contiguous")
    raise EmdError(f"emd_b200: {name} must be {dtype} (got {t.dtype})")


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)


def stream() -> int:
    """``cudaStream_t`` of torch's current stream on the current device (the raw-handle query: ``torch.cuda.current_stream()``
    builds a Python Stream object per call, ~15 us, and the step makes ~30 such calls)."""
    if _raw_stream is not None:
        return _raw_stream(torch.cuda.current_device())
    return torch.cuda.current_stream().cuda_stream


def launch_count() -> int:
    """Kernels launched by the library so far in this process."""
    return int(lib().emd_launch_count())


class profile:
    """``with _C.profile() as prof: ...`` brackets every library kernel with CUDA events on
    its own stream; ``prof.result()`` -> {kernel: (total_ms, launches)}."""

    def __enter__(self):
        L = lib()
        n = L.emd_kernel_id_count()
        self._ms = (ctypes.c_double * n)()
        self._cnt = (ctypes.c_longlong * n)()
        L.emd_profile_collect(self._ms, self._cnt)  # drop stale records
        for i in range(n):
            self._ms[i] = 0.0
            self._cnt[i] = 0
        L.emd_profile_enable(1)
        return self

    def __exit__(self, *exc):
        L = lib()
        L.emd_profile_enable(0)
        L.emd_profile_collect(self._ms, self._cnt)
        return False

    def result(self):
        L = lib()
        return {L.emd_kernel_name(i).decode(): (float(self._ms[i]), int(self._cnt[i]))
                for i in range(L.emd_kernel_id_count()) if self._cnt[i] > 0}
