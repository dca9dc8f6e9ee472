"""ctypes binding of libnrms_b200.so (C-ABI declared in include/nrms_b200.h).

There is NO fallback: if the shared library is missing or a call fails, an exception is
raised.  Build it with `python -c "import __graft_entry__ as g; g.build()"` (nvcc, sm_100a).
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libnrms_b200.so")

ABI_VERSION = 1


class NrmsError(RuntimeError):
    pass


class EncoderDims(C.Structure):
    """struct nrms_encoder_dims (include/nrms_b200.h)."""
    _fields_ = [
        ("n_seq", C.c_int32),
        ("seq_len", C.c_int32),
        ("d_model", C.c_int32),
        ("n_heads", C.c_int32),
        ("d_query", C.c_int32),
        ("vocab", C.c_int32),
        ("dropout_p", C.c_float),
        ("gemm_mode", C.c_int32),
        ("seed", C.c_uint64),
    ]


_P = C.c_void_p
_I32 = C.c_int32
_I64 = C.c_int64
_F = C.c_float
_DIMS = C.POINTER(EncoderDims)

# name -> (restype, argtypes); every symbol include/nrms_b200.h declares
SIGNATURES = {
    "nrms_abi_version": (C.c_int, []),
    "nrms_last_error": (C.c_char_p, []),
    "nrms_launch_count": (_I64, []),
    "nrms_profile_enable": (None, [C.c_int]),
    "nrms_profile_collect": (C.c_int, [C.c_char_p, _I64]),
    "nrms_encoder_param_count": (_I64, [_I32, _I32]),
    "nrms_encoder_saved_bytes": (_I64, [_DIMS]),
    "nrms_encoder_scratch_bytes": (_I64, [_DIMS]),
    "nrms_news_encoder_fwd": (C.c_int, [_DIMS, _P, _P, _P, _P, _P, _I64, _P]),
    "nrms_news_encoder_bwd": (C.c_int, [_DIMS, _P, _P, _P, _P, _P, _I64, _P, _I64, _P, _P, _P]),
    "nrms_news_encoder_bwd_phase": (C.c_int, [_DIMS, _P, _P, _P, _P, _P, _I64, _P, _I64, _P, _P, _I32, _P]),
    "nrms_user_encoder_bwd_phase": (C.c_int, [_DIMS, _P, _P, _P, _P, _I64, _P, _I64, _P, _P, _I32, _P]),
    "nrms_user_encoder_fwd": (C.c_int, [_DIMS, _P, _P, _P, _P, _I64, _P]),
    "nrms_user_encoder_fwd_gather": (C.c_int, [_DIMS, _P, _P, _P, _P, _P, _I64, _P]),
    "nrms_user_encoder_bwd": (C.c_int, [_DIMS, _P, _P, _P, _P, _I64, _P, _I64, _P, _P, _P]),
    "nrms_score_fwd": (C.c_int, [_I32, _I32, _I32, _P, _P, _P, _P, _P]),
    "nrms_score_cached": (C.c_int, [_I32, _I32, _I32, _P, _I64, _P, _P, _P, _P, _P]),
    "nrms_score_bwd": (C.c_int, [_I32, _I32, _I32, _P, _P, _P, _P, _P, _P, _P]),
    "nrms_score_ce_fwd_bwd": (C.c_int, [_I32, _I32, _I32, _I32, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "nrms_embedding_plan_bytes": (_I64, [_I64, _I32]),
    "nrms_embedding_plan": (C.c_int, [_P, _I64, _I32, _P, _I64, _P]),
    "nrms_embedding_grad_dense": (C.c_int, [_P, _I64, _P, _I64, _I32, _I32, _P, _P]),
    "nrms_embedding_plan_unique": (C.c_int, [_P, _I64, _I32, _P, _P]),
    "nrms_adam_step": (C.c_int, [_P, _P, _P, _P, _I64, _I32, _F, _F, _F, _F, _F, _P]),
    "nrms_rank_metrics": (C.c_int, [_P, _P, _P, _I64, _I32, _P, _P]),
    "nrms_rank_metrics_padded": (C.c_int, [_P, _I64, _P, _P, _I64, _I32, _P, _P]),
    "nrms_rank_metrics_rows": (C.c_int, [_P, _I64, _P, _I64, _P, _I64, _I32, _P, _P]),
    "nrms_assemble_batch": (C.c_int, [_P, _I32, _P, _P, _P, _P, _P, _I64, _I32, _I32, _I32, _P, _P, _P, _P, _P, _P, _P, _P]),
    "nrms_rank_positions": (C.c_int, [_P, _I64, _P, _I64, _P, _P]),
    "nrms_gather_rows_f32": (C.c_int, [_P, _I64, _I32, _P, _I64, _I64, _P, _P]),
    "nrms_gather_rows_i64": (C.c_int, [_P, _I64, _I32, _P, _I64, _I64, _P, _P]),
    "nrms_dropout_mask": (C.c_int, [C.c_uint64, C.c_uint32, _F, _I64, _I32, _P, _P]),
    "nrms_gemm_selftest_bytes": (_I64, [_I32, _I32, _I32, _I32]),
    "nrms_gemm_selftest": (C.c_int, [_I32, _P, _P, _P, _I32, _I32, _I32, _P, _I64, _P]),
    "nrms_validate_ids": (C.c_int, [_P, _I64, _I64, _P, _P]),
    # the `nrms` sibling variant (model/nrms.py)
    "nrms_linear_work_bytes": (_I64, [_I32, _I32, _I32]),
    "nrms_linear_fwd": (C.c_int, [_P, _P, _P, _P, _I32, _I32, _I32, _P, _I64, _P]),
    "nrms_linear_bwd": (C.c_int, [_P, _P, _P, _P, _P, _P, _I32, _I32, _I32, _P, _I64, _P]),
    "nrms_dropout_apply": (C.c_int, [C.c_uint64, C.c_uint32, _F, _I64, _I32, _P, _P, _P]),
    "nrms_masked_attention_fwd": (C.c_int, [_P, _P, _I32, _I32, _I32, _I32, _F, C.c_uint64, _P, _P, _P]),
    "nrms_masked_attention_bwd": (C.c_int, [_P, _P, _P, _P, _I32, _I32, _I32, _I32, _F, C.c_uint64, _P, _P]),
    "nrms_masked_pool_fwd": (C.c_int, [_P, _P, _P, _P, _I32, _I32, _I32, _I32, _P, _P, _P]),
    "nrms_masked_pool_bwd": (C.c_int, [_P, _P, _P, _P, _P, _P, _I32, _I32, _I32, _I32, _P, _P, _P, _P, _P]),
}

_lib: Optional[C.CDLL] = None


def load() -> C.CDLL:
    """Load the library and bind every declared symbol; raises NrmsError when unavailable."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise NrmsError(
            f"{LIB_PATH} is missing: the CUDA extension has not been built. Run "
            "`python -c 'import __graft_entry__ as g; g.build()'` (needs nvcc). "
            "There is no CPU/PyTorch fallback for this path.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        try:
            fn = getattr(lib, name)
        except AttributeError as e:
            raise NrmsError(f"{LIB_PATH} does not export {name}") from e
        fn.restype = res
        fn.argtypes = args
    v = lib.nrms_abi_version()
    if v != ABI_VERSION:
        raise NrmsError(f"ABI version mismatch: library {v}, binding {ABI_VERSION}")
    _lib = lib
    return lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().nrms_last_error()
        raise NrmsError(f"{what} failed with code {rc}: {msg.decode() if msg else ''}")


def ptr(t) -> Optional[int]:
    """Device pointer of a torch tensor (None -> NULL)."""
    if t is None:
        return None
    return t.data_ptr()
