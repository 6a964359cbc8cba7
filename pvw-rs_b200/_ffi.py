"""ctypes binding of the C ABI declared in include/pvw_b200.h (libpvw_b200.so).

This is the same boundary a Rust shim of pvw-rs binds.  There is no CPU fallback: if the
library is missing or no CUDA device exists, loading / context creation fails loudly.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libpvw_b200.so")

PVW_OK = 0
PVW_IO_HOST, PVW_IO_DEVICE = 0, 1
PVW_ENC_C1_ONLY, PVW_ENC_C2_ONLY, PVW_ENC_PUSH_C1 = 2, 4, 8
PVW_IN_SECRET_I8, PVW_IN_ERROR_I32, PVW_IN_ERROR_I16 = 0x100, 0x200, 0x400
STATUS_NAMES = {
    0: "Ok", -1: "InvalidParameters", -2: "DimensionMismatch", -3: "IndexOutOfBounds", -4: "EncryptionError",
    -5: "DecryptionError", -6: "KeyGenerationError", -7: "InternalError", -8: "DeserializationError", -9: "InsufficientData",
}


class PvwParamsDesc(C.Structure):
    _fields_ = [("n", C.c_uint32), ("k", C.c_uint32), ("ell", C.c_uint32), ("L", C.c_uint32),
                ("moduli", C.POINTER(C.c_uint64)), ("psi", C.POINTER(C.c_uint64)),
                ("secret_variance", C.c_float), ("error_bound_1", C.c_uint64), ("error_bound_2", C.c_uint64),
                ("row0", C.c_uint32), ("nrows", C.c_uint32), ("device", C.c_int32)]


class PvwWireLayout(C.Structure):
    _fields_ = [(name, C.c_uint64) for name in ("poly_bytes", "record_bytes", "params_bytes", "pk_row_bytes", "ciphertext_bytes",
                                                "crs_bytes", "ct_c1_offset", "ct_c2_offset", "ct_params_offset")]


class PvwShardHandle(C.Structure):
    _fields_ = [("bytes", C.c_uint8 * 192)]


_vp, _u32, _u64, _i64 = C.c_void_p, C.c_uint32, C.c_uint64, C.c_int64

# name -> (restype, argtypes): every symbol include/pvw_b200.h declares
SIGNATURES = {
    "pvw_ctx_create": (C.c_int, [C.POINTER(_vp), C.POINTER(PvwParamsDesc)]),
    "pvw_ctx_destroy": (None, [_vp]),
    "pvw_last_error": (C.c_char_p, [_vp]),
    "pvw_params_bigint": (C.c_int, [_vp, C.c_int, _vp, _u32, C.POINTER(_u32)]),
    "pvw_params_psi": (C.c_int, [_vp, _vp]),
    "pvw_params_correctness_condition": (C.c_int, [_vp, C.POINTER(C.c_int)]),
    "pvw_crs_upload": (C.c_int, [_vp, _vp, _u32]),
    "pvw_crs_download": (C.c_int, [_vp, _vp]),
    "pvw_crs_generate_deterministic": (C.c_int, [_vp, _vp, _vp]),
    "pvw_crs_generate_from_tag": (C.c_int, [_vp, C.c_char_p, _vp]),
    "pvw_crs_expand_seed": (C.c_int, [_u32, _u32, _u32, _vp, _vp, _vp]),
    "pvw_crs_tag_to_seed": (C.c_int, [C.c_char_p, _vp]),
    "pvw_pk_upload_rows": (C.c_int, [_vp, _u32, _u32, _vp, _u32]),
    "pvw_pk_download_rows": (C.c_int, [_vp, _u32, _u32, _vp]),
    "pvw_pk_num_keys": (C.c_int, [_vp, C.POINTER(_u32)]),
    "pvw_keygen_batch": (C.c_int, [_vp, _u32, _u32, _vp, _vp, _u32]),
    "pvw_crs_multiply_by_randomness": (C.c_int, [_vp, _u32, _vp, _vp]),
    "pvw_ct_reserve": (C.c_int, [_vp, _u32]),
    "pvw_encrypt_batch": (C.c_int, [_vp, _u32, _u32, _u32, _u32, _vp, _vp, _vp, _vp, _u32]),
    "pvw_ct_download": (C.c_int, [_vp, _u32, _vp, _vp]),
    "pvw_ct_upload": (C.c_int, [_vp, _u32, _vp, _vp]),
    "pvw_ct_c1_device_ptr": (C.c_int, [_vp, _u32, C.POINTER(_vp), C.POINTER(_u64)]),
    "pvw_decrypt_batch": (C.c_int, [_vp, _u32, _vp, _u32, _vp, _vp, _vp, _u32]),
    "pvw_decode_batch": (C.c_int, [_vp, _u32, _vp, _vp]),
    "pvw_ntt_forward_small": (C.c_int, [_vp, _u32, _vp, _vp]),
    "pvw_encode_scalars": (C.c_int, [_vp, _u32, _vp, _vp]),
    "pvw_wire_layout_get": (C.c_int, [_vp, C.POINTER(PvwWireLayout)]),
    "pvw_wire_params": (C.c_int, [_vp, _vp, _u64]),
    "pvw_wire_ct_serialize": (C.c_int, [_vp, _u32, _u32, _vp, _u64, _u32]),
    "pvw_wire_ct_deserialize": (C.c_int, [_vp, _u32, _u32, _vp, _u64, _u32]),
    "pvw_wire_pk_serialize_rows": (C.c_int, [_vp, _u32, _u32, _vp, _u32]),
    "pvw_wire_pk_deserialize_rows": (C.c_int, [_vp, _u32, _u32, _vp, _u32]),
    "pvw_wire_crs_serialize": (C.c_int, [_vp, _vp, _u64, _u32]),
    "pvw_wire_crs_deserialize": (C.c_int, [_vp, _vp, _u64, _u32]),
    "pvw_wire_polys_serialize": (C.c_int, [_vp, _u32, _vp, _vp]),
    "pvw_wire_polys_deserialize": (C.c_int, [_vp, _u32, _vp, _vp]),
    "pvw_shard_export": (C.c_int, [_vp, _u32, C.POINTER(PvwShardHandle)]),
    "pvw_shard_connect": (C.c_int, [_vp, _u32, _u32, C.POINTER(PvwShardHandle)]),
    "pvw_shard_push_c1": (C.c_int, [_vp, _u32, _u32]),
    "pvw_shard_wait_c1": (C.c_int, [_vp]),
    "pvw_shard_release_c1": (C.c_int, [_vp]),
    "pvw_shard_disconnect": (C.c_int, [_vp]),
    "pvw_ctx_synchronize": (C.c_int, [_vp]),
    "pvw_ctx_stream": (_vp, [_vp]),
    "pvw_ctx_set_option": (C.c_int, [_vp, C.c_char_p, _i64]),
    "pvw_ctx_profile": (C.c_int, [_vp, C.c_int, C.POINTER(C.c_double), C.POINTER(_u64), C.POINTER(C.c_double)]),
    "pvw_ctx_launch_count": (_u64, [_vp]),
    "pvw_version": (C.c_char_p, []),
}

KERNEL_KINDS = ["ntt_small", "mac_gemm", "decode_rns", "crt_lift", "decode_tail", "permute", "wire", "expand", "decode_fused", "imma_gemm", "ntt_planes"]

_lib = None


def load(path: str | None = None) -> C.CDLL:
    """dlopen libpvw_b200.so and declare every entry point.  Raises if the library has not been built."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or LIB_PATH
    if not os.path.exists(p):
        raise RuntimeError(f"{p} is missing: build it with `python pvw-rs_b200/build.py` (there is no CPU fallback)")
    lib = C.CDLL(p)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)           # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    if path is None:
        _lib = lib
    return lib
