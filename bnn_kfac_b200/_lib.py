"""ctypes binding of libbk_kfac.so (include/bk_kfac.h).

There is no CPU fallback: if the shared library is missing the import of any compute entry point
raises, and every call checks its return code.
"""
from __future__ import annotations

import ctypes as C
import threading
from pathlib import Path

from . import _build

BK_PREC_FP32 = 0
BK_PREC_BF16 = 1
BK_PREC_BF16X3 = 3
BK_SMALL_D_MAX = 176
BK_SMALL64_MAX_DIM = 112
BK_SMALL64_MAX_ELEMS = 12544
BK_SMALL64_MAX_BATCH = 16
BK_BLOCK_INV_MAX_DIM = 160

SYRK_LOWER_ONLY = 1
SYRK_NO_OVERLAP = 2
SYRK_STAGE_PERSISTENT = 8
SYRK_ROW_MAJOR = 4

GEMM_SYRK_LOWER = 1
GEMM_MIRROR = 2
GEMM_TRI_A = 4
GEMM_TRI_B = 8
GEMM_RELU = 16
GEMM_TRI_B_UPPER = 32


def gemm_tri_koff(koff: int) -> int:
    """Flag bits for BK_GEMM_TRI_KOFF(koff) (include/bk_kfac.h)."""
    assert koff % 8 == 0
    return (koff // 8) << 8

_ERRORS = {
    -2: "BK_ERR_ARG (bad argument or alignment)",
    -3: "BK_ERR_DRIVER (cuTensorMapEncodeTiled unavailable)",
    -4: "BK_ERR_TMAP (tensor-map encode failed)",
    -5: "BK_ERR_CUDA (launch/runtime error)",
    -6: "BK_ERR_WORKSPACE (workspace too small or misaligned)",
    -7: "BK_ERR_ARCH (device is not compute capability 10.x)",
}

_p, _ll, _i, _f, _u, _ull, _sz = (C.c_void_p, C.c_longlong, C.c_int, C.c_float, C.c_uint,
                                  C.c_ulonglong, C.c_size_t)

# name -> (restype, argtypes); mirrors include/bk_kfac.h declaration by declaration
SIGNATURES = {
    "bk_version": (C.c_char_p, []),
    "bk_device_check": (_i, []),
    "bk_launch_count": (_ull, []),
    "bk_set_cta_group": (None, [_i]),
    "bk_set_syrk_tuning": (None, [_i]),
    "bk_set_chol_graph": (None, [_i]),
    "bk_set_conv_fast": (None, [_i]),
    "bk_set_chol_far_sms": (None, [_i]),
    "bk_set_chol_lookahead": (None, [_i]),
    "bk_set_eigh_mode": (None, [_i]),
    "bk_set_eigh_pair_width": (None, [_i]),
    "bk_gemm_nt": (_i, [_p, _p, _ll, _ll, _p, _p, _ll, _ll, _i, _i, _i, _i, _i, _i, _f, _f,
                        _p, _ll, _ll, _p, _ll, _p, _p, _ll, _ll, _p]),
    "bk_transpose_split": (_i, [_p, _ll, _i, _i, _f, _i, _p, _p, _ll, _p]),
    "bk_convert_split": (_i, [_p, _ll, _i, _i, _f, _i, _p, _p, _ll, _p]),
    "bk_philox_normal": (_i, [_ull, _u, _u, _i, _i, _i, _p, _ll, _ll, _p, _p, _ll, _ll, _p]),
    "bk_syrk_workspace_bytes": (_sz, [_i, _i, _i, _i]),
    "bk_syrk_accum": (_i, [_p, _ll, _p, _ll, _i, _i, _i, _f, _f, _f, _i, _p, _sz, _p]),
    "bk_syrk_grouped_workspace_bytes": (_sz, [C.POINTER(_i), C.POINTER(_i), C.POINTER(_i), _i, _i]),
    "bk_syrk_accum_grouped": (_i, [C.POINTER(_p), C.POINTER(_ll), C.POINTER(_p), C.POINTER(_i), C.POINTER(_ll),
                                   C.POINTER(_i), C.POINTER(_i), C.POINTER(_i), C.POINTER(_f),
                                   C.POINTER(_f), C.POINTER(_f), _i, _i, _i, _p, _sz, _p]),
    "bk_sym_finalize": (_i, [C.POINTER(_p), C.POINTER(_ll), C.POINTER(_i), _i, _f, _p]),
    "bk_syrk_accum_staged_grouped": (_i, [C.POINTER(_p), C.POINTER(_ll), C.POINTER(_p), C.POINTER(_p),
                                          C.POINTER(_ll), C.POINTER(_i), C.POINTER(_i), C.POINTER(_f),
                                          C.POINTER(_f), _i, _i, _i, _p]),
    "bk_syrk_accum_staged": (_i, [_p, _ll, _p, _p, _ll, _i, _i, _f, _f, _i, _p]),
    "bk_conv_a_accum": (_i, [_p, _ll, _p, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _f, _f, _p]),
    "bk_im2col_split": (_i, [_p] + [_i] * 10 + [_f, _i, _p, _p, _ll, _p]),
    "bk_conv_g_accum": (_i, [_p, _ll, _p, _i, _i, _i, _f, _f, _f, _p]),
    "bk_diag_accum": (_i, [_p, _p, _p, _i, _i, _f, _f, _p]),
    "bk_diag_invert": (_i, [_p, _p, _ll, _f, _f, _p]),
    "bk_diag_sample": (_i, [_p, _p, _ll, _i, _ull, _u, _u, _p, _p]),
    "bk_diag_quadform": (_i, [_p, _p, _ll, _p, _ll, _i, _p]),
    "bk_sample_to_weights": (_i, [_p, _p, _p, _i, _i, _i, _i, _p, _p, _p, _ll, _p, _p]),
    "bk_conv2d_relu_pool": (_i, [_p, _ll, _p, _p, _p] + [_i] * 14 + [_p]),
    "bk_predictive_moments": (_i, [_p, _i, _i, _i, _i, _p, _p, _p]),
    "bk_frob_dot": (_i, [_p, _p, _ll, _p, _ll, _ll, _i, _i, _i, _p]),
    "bk_spd_inverse_f64": (_i, [C.POINTER(_p), C.POINTER(_ll), C.POINTER(_i), C.POINTER(C.c_double),
                                C.POINTER(C.c_double), C.POINTER(_p), _i, _p, _p]),
    "bk_kron_quadform_f64": (_i, [_p, _ll, _i, _i, _i, _p, _p, _p, _i, _p]),
    "bk_eigh_workspace_bytes": (_sz, [C.POINTER(_i), _i]),
    "bk_eigh_batched": (_i, [C.POINTER(_p), C.POINTER(_ll), C.POINTER(_p), C.POINTER(_p), C.POINTER(_i),
                             _i, _f, _i, _p, _sz, _p]),
    "bk_ger_accum": (_i, [_p, _ll, _p, _i, _f, _f, _p]),
    "bk_kron": (_i, [_p, _i, _i, _p, _i, _i, _p, _p]),
    "bk_dominance": (_i, [_p, _ll, _i, _f, _p, _p, _i, _p, _p]),
    "bk_chol_trinv_f64": (_i, [_p, _ll, _i, C.c_double, _p, _ll, _p, _p]),
    "bk_band_mask": (_i, [_p, _ll, _i, _f, _i, _p, _p, _p, _ll, _p]),
    "bk_block_inverse": (_i, [_p, _ll, _i, _p, _p, _i, _i, C.c_double, _p, _ll, _i, _p, _p]),
    "bk_dominance_rows": (_i, [_p, _ll, _i, _i, _i, _f, _p, _p, _i, _p, _p]),
    "bk_tri_pack": (_i, [C.POINTER(_p), C.POINTER(_ll), C.POINTER(_i), _i, _p, _p]),
    "bk_tri_unpack": (_i, [C.POINTER(_p), C.POINTER(_ll), C.POINTER(_i), _i, _p, _f, _i, _p]),
    "bk_peer_alloc": (_i, [_sz, C.POINTER(_p)]),
    "bk_peer_free": (_i, [_p]),
    "bk_peer_export": (_i, [_p, _p]),
    "bk_peer_open": (_i, [_p, C.POINTER(_p)]),
    "bk_peer_close": (_i, [_p]),
    "bk_peer_read_u32": (_i, [_p, C.POINTER(C.c_uint)]),
    "bk_tile_packed_floats": (_ll, [C.POINTER(_i), _i]),
    "bk_tile_pack": (_i, [C.POINTER(_p), C.POINTER(_ll), C.POINTER(_i), C.POINTER(_ll), _i, _p, _p]),
    "bk_tile_pack_to": (_i, [C.POINTER(_p), C.POINTER(_ll), C.POINTER(_i), C.POINTER(_p), _i, _p]),
    "bk_peer_tile_unpack": (_i, [C.POINTER(_p), C.POINTER(_ll), C.POINTER(_i), _i, C.POINTER(_p), _i, _f, _i, _p]),
    "bk_peer_copy": (_i, [_p, _p, _ll, _i, _i, _p]),
    "bk_peer_signal": (_i, [C.POINTER(_p), _i, _i, C.c_uint, _p]),
    "bk_peer_wait": (_i, [_p, _i, C.c_uint, C.c_double, _p, _p]),
    "bk_inf_regularise": (_i, [_p, _ll, _p, _ll, _f, _f, _p, _p, _p]),
    "bk_inf_presample_workspace_bytes": (_sz, [_i, _i, _i, _i]),
    "bk_inf_presample": (_i, [_p, _ll, _i, _i, _p, _ll, _i, _i, _p, _p, _p, _p, _sz, _p]),
    "bk_inf_combine": (_i, [_p, _p, _p, _p, _ll, _p]),
    "bk_calibration_rows": (_i, [_p, _ll, _p, _i, _i, _p, _p, _p, _p, _p, _p, _p]),
    "bk_binned_stats": (_i, [_p, _p, _p, _ll, _p, _i, _i, _p, _p]),
    "bk_chol_inv_workspace_bytes": (_sz, [C.POINTER(_i), _i]),
    "bk_damp_chol_inv_batched": (_i, [C.POINTER(_p), C.POINTER(_p), C.POINTER(_i),
                                      C.POINTER(_f), C.POINTER(_f), _i, _p, _sz, _p]),
}

_lock = threading.Lock()
_lib = None


class BkError(RuntimeError):
    """A libbk_kfac entry point returned a negative status."""


def lib_path() -> Path:
    return _build.LIB_PATH


def load(build_if_missing: bool = True, strict: bool = True) -> C.CDLL:
    """Load (building first if needed) the kernel library. Raises if it cannot be had.

    strict=False (bring-up tools only) tolerates declared-but-missing symbols."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        path = lib_path()
        if build_if_missing:
            # no-op when the source digest matches the stamp; rebuilds after any edit of csrc/ or the header,
            # so a stale library can never be loaded silently against newer Python glue
            _build.build()
        elif not path.exists():
            raise BkError(f"{path} is missing: run `python -m bnn_kfac_b200._build`")
        elif not _build.is_current():
            raise BkError(f"{path} is stale (csrc/ or include/bk_kfac.h changed since it was built): run "
                          f"`python -m bnn_kfac_b200._build`")
        lib = C.CDLL(str(path))
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name, None)
            if fn is None:
                if strict:
                    raise BkError(f"{path} does not export {name} (declared in include/bk_kfac.h)")
                continue
            fn.restype = res
            fn.argtypes = args
        _lib = lib
        return lib


def check(rc: int, what: str) -> int:
    if rc < 0:
        raise BkError(f"{what} failed: {_ERRORS.get(rc, rc)}")
    return rc


def require_device() -> None:
    """Fail loudly unless the current CUDA device can run the sm_100a kernels."""
    import torch
    if not torch.cuda.is_available():
        raise BkError("bnn_kfac_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    check(load().bk_device_check(), "bk_device_check")


def ptr(t) -> int:
    """Device pointer of a torch tensor (or 0 for None)."""
    return 0 if t is None else t.data_ptr()


def stream_ptr() -> int:
    import torch
    return torch.cuda.current_stream().cuda_stream


def nvtx_range(name: str):
    """Decorator: wraps a stage entry point in an NVTX range `bk/<name>` (one range per stage
    of the path - factor update, inversion, sampling, predictive - so that an Nsight Systems / ncu --nvtx timeline
    of a user's run shows them by name).  A push/pop pair costs ~100 ns; no-op without CUDA."""
    import functools

    def deco(fn):
        @functools.wraps(fn)
        def wrapped(*args, **kwargs):
            import torch
            if not torch.cuda.is_available():
                return fn(*args, **kwargs)
            torch.cuda.nvtx.range_push("bk/" + name)
            try:
                return fn(*args, **kwargs)
            finally:
                torch.cuda.nvtx.range_pop()
        return wrapped
    return deco
