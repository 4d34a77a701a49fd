"""ctypes binding of libcvcs_b200.so (the C-ABI declared in include/cvcs_b200.h).

There is no fallback: if the shared library is missing, importing this module raises.  The
library itself is plain CUDA C (no torch symbols); tensors cross the boundary as raw device
pointers + the current CUDA stream handle.
"""
from __future__ import annotations

import ctypes as C
import os

_PKG = os.path.dirname(os.path.abspath(__file__))
# CVCS_B200_LIB selects an experiment build (python -m cvcs_b200.build --tag X -D...) for A/B timing runs
LIB_PATH = os.environ.get("CVCS_B200_LIB") or os.path.join(_PKG, "libcvcs_b200.so")

# dtype / layout tags (enum cvcs_dtype / cvcs_layout)
F32, BF16, U8, I64, I32 = 0, 1, 2, 3, 4
NCHW, NHWC = 0, 1

OK, ERR_INVALID_ARG, ERR_UNSUPPORTED, ERR_CUDA, ERR_WORKSPACE = 0, -1, -2, -3, -4


class CvcsError(RuntimeError):
    """A C-ABI call returned a negative status."""

    def __init__(self, code: int, message: str):
        super().__init__(f"cvcs_b200 error {code}: {message}")
        self.code = code
        self.message = message


def _load() -> C.CDLL:
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: build it with `python -m cvcs_b200.build` "
            "(cvcs_b200 has no CPU or PyTorch fallback)")
    return C.CDLL(LIB_PATH)


lib = _load()

_vp, _i, _ll, _d, _sz = C.c_void_p, C.c_int, C.c_longlong, C.c_double, C.c_size_t

# name -> (restype, argtypes); mirrors include/cvcs_b200.h one to one
SIGNATURES = {
    "cvcs_abi_version": (_i, []),
    "cvcs_last_error": (C.c_char_p, []),
    "cvcs_sm_count": (_i, []),
    "cvcs_workspace_bytes": (_sz, []),
    "cvcs_stream_capture_id": (_i, [_vp, C.POINTER(C.c_ulonglong)]),
    "cvcs_set_option": (_i, [_i, _i]),
    "cvcs_label_hist": (_i, [_vp, _i, _ll, _i, _ll, _vp, _vp, _vp, _vp, _vp]),
    "cvcs_total_weight": (_i, [_vp, _vp, _i, _ll, _vp, _vp]),
    "cvcs_labels_prepare": (_i, [_vp, _ll, _i, _ll, _vp, _vp, _vp, _vp, _vp]),
    "cvcs_ce_fused": (_i, [_vp, _i, _i, _vp, _i, _vp, _ll, _i, _i, _i, _i, _d, _vp, _vp, _vp, _i, _vp, _vp,
                           _vp, _vp, _vp]),
    "cvcs_ce_fused_tw": (_i, [_vp, _i, _i, _vp, _i, _vp, _ll, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _ll, _vp, _vp, _vp, _i,
                              _vp, _vp, _vp, _vp, _vp]),
    "cvcs_xchg_create": (_i, [C.POINTER(_vp), _i, _i]),
    "cvcs_xchg_local_handle": (_i, [_vp, _vp]),
    "cvcs_xchg_open_peer": (_i, [_vp, _i, _vp]),
    "cvcs_xchg_set_peer": (_i, [_vp, _i, _vp]),
    "cvcs_xchg_local_block": (_vp, [_vp]),
    "cvcs_xchg_state": (_i, [_vp, C.POINTER(C.c_ulonglong), C.POINTER(C.c_ulonglong)]),
    "cvcs_xchg_poke": (_i, [_vp, _i, C.c_ulonglong, _d, _vp]),
    "cvcs_xchg_allreduce_f64": (_i, [_vp, _vp, _i, _vp]),
    "cvcs_xchg_destroy": (_i, [_vp]),
    "cvcs_eval_fused": (_i, [_vp, _i, _i, _vp, _i, _ll, _i, _i, _i, _i, _vp, _i, _vp, _vp, _vp, _vp]),
    "cvcs_scale_inplace": (_i, [_vp, _i, _ll, _vp, _vp]),
    "cvcs_argmax": (_i, [_vp, _i, _i, _i, _i, _i, _i, _vp, _i, _vp]),
    "cvcs_confmat": (_i, [_vp, _i, _vp, _i, _ll, _i, _ll, _vp, _vp, _vp, _vp]),
    "cvcs_tile_normalize": (_i, [_vp, _i, _i, _i, _vp, _vp, _i, _i, _i, _vp, _vp, _vp, _i, _vp, _vp, _i, _vp, _i, _ll,
                                 _vp, _vp]),
    "cvcs_tile_context": (_i, [_vp, _i, _i, _i, _vp, _vp, _i, _i, _vp, _vp]),
    "cvcs_vote": (_i, [_vp, _i, _i, _ll, _i, _vp, _i, _vp]),
    "cvcs_colorize": (_i, [_vp, _i, _ll, _vp, _i, _vp, _vp]),
    "cvcs_stitch": (_i, [_vp, _i, _i, _i, _vp, _i, _i, _vp, _i, _i, _vp]),
    "cvcs_host_ctx_create": (_i, [C.POINTER(_vp), _i, _ll, _i, _i]),
    "cvcs_host_ctx_destroy": (_i, [_vp]),
    "cvcs_host_ce_fused": (_i, [_vp, _vp, _i, _i, _vp, _i, _vp, _ll, _i, _i, _i, _i, _i, _vp, _vp, _i, _vp, _vp,
                                _vp]),
    "cvcs_host_ctx_device_ptr": (_vp, [_vp, _i]),
}

for _name, (_res, _args) in SIGNATURES.items():
    _fn = getattr(lib, _name)  # AttributeError if the .so does not export a declared symbol
    _fn.restype = _res
    _fn.argtypes = _args


def last_error() -> str:
    return lib.cvcs_last_error().decode("utf-8", "replace")


def check(code: int) -> None:
    if code != 0:
        raise CvcsError(code, last_error())


OPT_CE_PATH, OPT_TMA_STAGES, OPT_TMA_WAIT_HINT, OPT_TMA_VECP, OPT_TMA_CTAS, OPT_TILE_CTAS, OPT_RESERVE_SMS, OPT_PDL, OPT_L2_HINT = 0, 1, 2, 3, 4, 5, 6, 7, 8
CE_PATH_AUTO, CE_PATH_TMA, CE_PATH_DIRECT, CE_PATH_GENERIC = 0, 1, 2, 3


def set_option(option: int, value: int) -> None:
    check(lib.cvcs_set_option(option, value))


def abi_version() -> int:
    return lib.cvcs_abi_version()


def workspace_bytes() -> int:
    return int(lib.cvcs_workspace_bytes())


def stream_capture_id(stream: int) -> int:
    """Sequence number of the CUDA-graph capture the stream belongs to (0: not capturing)."""
    out = C.c_ulonglong(0)
    check(lib.cvcs_stream_capture_id(C.c_void_p(stream), C.byref(out)))
    return int(out.value)
