"""pvw-rs_b200 -- B200-native implementation of the pvw-rs hot path (multi-receiver PVW encryption and per-party
decryption over R_q in RNS/NTT form).  Hand-written CUDA kernels for sm_100a behind the C ABI of
include/pvw_b200.h; this package is the host-side mirror of the crate's API surface for that path
(src/lib.rs:31-55).  There is no CPU fallback: everything below needs libpvw_b200.so and a CUDA device."""
from . import _ffi, sharding
from .errors import PvwError
from .engine import Engine
from .api import (GlobalPublicKey, Party, PublicKey, PvwCiphertext, PvwCrs, PvwParameters, PvwParametersBuilder,
                  SecretKey, decrypt_party_shares, decrypt_party_value, encrypt, encrypt_all_party_shares,
                  encrypt_broadcast, encrypt_party_shares, sample_uniform_coefficients, sample_vec_cbd)

from . import serde  # noqa: E402  (needs api)

__all__ = ["Engine", "sharding", "serde", "PvwError", "GlobalPublicKey", "Party", "PublicKey", "PvwCiphertext", "PvwCrs", "PvwParameters",
           "PvwParametersBuilder", "SecretKey", "decrypt_party_shares", "decrypt_party_value", "encrypt",
           "encrypt_all_party_shares", "encrypt_broadcast", "encrypt_party_shares", "sample_uniform_coefficients",
           "sample_vec_cbd"]
