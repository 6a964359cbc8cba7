//! pvw-b200-sys -- `extern "C"` declarations of include/pvw_b200.h, one per entry point (generated from the header by the
//! script in rust/README.md's history; bindgen emits the same).  See the header for what each call replaces in pvw-rs.
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_int, c_void};

#[repr(C)]
pub struct pvw_ctx { _private: [u8; 0] }

/// pvw_status
pub const PVW_OK: c_int = 0;
pub const PVW_ERR_INVALID_PARAMETERS: c_int = -1;
pub const PVW_ERR_DIMENSION_MISMATCH: c_int = -2;
pub const PVW_ERR_INDEX_OUT_OF_BOUNDS: c_int = -3;
pub const PVW_ERR_ENCRYPTION: c_int = -4;
pub const PVW_ERR_DECRYPTION: c_int = -5;
pub const PVW_ERR_KEYGEN: c_int = -6;
pub const PVW_ERR_INTERNAL: c_int = -7;
pub const PVW_ERR_DESERIALIZATION: c_int = -8;
pub const PVW_ERR_INSUFFICIENT_DATA: c_int = -9;

/// flags
pub const PVW_IO_HOST: u32 = 0;
pub const PVW_IO_DEVICE: u32 = 1;
pub const PVW_ENC_C1_ONLY: u32 = 2;
pub const PVW_ENC_C2_ONLY: u32 = 4;
pub const PVW_ENC_PUSH_C1: u32 = 8;
pub const PVW_IN_SECRET_I8: u32 = 0x100;
pub const PVW_IN_ERROR_I32: u32 = 0x200;
pub const PVW_IN_ERROR_I16: u32 = 0x400;

#[repr(C)]
pub struct pvw_params_desc {
    pub n: u32, pub k: u32, pub ell: u32, pub l_count: u32,
    pub moduli: *const u64, pub psi: *const u64,
    pub secret_variance: f32, pub error_bound_1: u64, pub error_bound_2: u64,
    pub row0: u32, pub nrows: u32, pub device: i32,
}
#[repr(C)]
#[derive(Default, Clone, Copy)]
pub struct pvw_wire_layout {
    pub poly_bytes: u64, pub record_bytes: u64, pub params_bytes: u64, pub pk_row_bytes: u64, pub ciphertext_bytes: u64,
    pub crs_bytes: u64, pub ct_c1_offset: u64, pub ct_c2_offset: u64, pub ct_params_offset: u64,
}
#[repr(C)]
#[derive(Clone, Copy)]
pub struct pvw_shard_handle { pub bytes: [u8; 192] }

extern "C" {
    pub fn pvw_ctx_create(out_: *mut *mut pvw_ctx, desc: *const pvw_params_desc) -> c_int;
    pub fn pvw_ctx_destroy(ctx: *mut pvw_ctx);
    pub fn pvw_last_error(ctx: *const pvw_ctx) -> *const c_char;
    pub fn pvw_params_bigint(ctx: *const pvw_ctx, which: c_int, out_: *mut u64, cap: u32, nwords: *mut u32) -> c_int;
    pub fn pvw_params_psi(ctx: *const pvw_ctx, out_: *mut u64) -> c_int;
    pub fn pvw_params_correctness_condition(ctx: *const pvw_ctx, ok: *mut c_int) -> c_int;
    pub fn pvw_crs_upload(ctx: *mut pvw_ctx, A: *const u64, flags: u32) -> c_int;
    pub fn pvw_crs_download(ctx: *mut pvw_ctx, A: *mut u64) -> c_int;
    pub fn pvw_crs_generate_deterministic(ctx: *mut pvw_ctx, seed: *const u8, A_out: *mut u64) -> c_int;
    pub fn pvw_crs_generate_from_tag(ctx: *mut pvw_ctx, tag: *const c_char, A_out: *mut u64) -> c_int;
    pub fn pvw_crs_expand_seed(k: u32, ell: u32, L: u32, moduli: *const u64, seed: *const u8, A_out: *mut u64) -> c_int;
    pub fn pvw_crs_tag_to_seed(tag: *const c_char, seed_out: *mut u8) -> c_int;
    pub fn pvw_pk_upload_rows(ctx: *mut pvw_ctx, row: u32, count: u32, B: *const u64, flags: u32) -> c_int;
    pub fn pvw_pk_download_rows(ctx: *mut pvw_ctx, row: u32, count: u32, B: *mut u64) -> c_int;
    pub fn pvw_pk_num_keys(ctx: *const pvw_ctx, num_keys: *mut u32) -> c_int;
    pub fn pvw_keygen_batch(ctx: *mut pvw_ctx, row: u32, count: u32, sk: *const c_void, e: *const c_void, flags: u32) -> c_int;
    pub fn pvw_crs_multiply_by_randomness(ctx: *mut pvw_ctx, D: u32, r_hat: *const u64, out_: *mut u64) -> c_int;
    pub fn pvw_ct_reserve(ctx: *mut pvw_ctx, capacity: u32) -> c_int;
    pub fn pvw_encrypt_batch(ctx: *mut pvw_ctx, slot0: u32, D: u32, c1_lo: u32, c1_hi: u32, m: *const u64, r: *const c_void, e1: *const c_void, e2: *const c_void, flags: u32) -> c_int;
    pub fn pvw_ct_download(ctx: *mut pvw_ctx, slot: u32, c1: *mut u64, c2: *mut u64) -> c_int;
    pub fn pvw_ct_upload(ctx: *mut pvw_ctx, slot: u32, c1: *const u64, c2: *const u64) -> c_int;
    pub fn pvw_ct_c1_device_ptr(ctx: *mut pvw_ctx, slot: u32, ptr: *mut *mut c_void, slot_stride: *mut u64) -> c_int;
    pub fn pvw_decrypt_batch(ctx: *mut pvw_ctx, D: u32, dealer_slots: *const u32, P: u32, party_idx: *const u32, sk: *const c_void, out_: *mut u64, flags: u32) -> c_int;
    pub fn pvw_decode_batch(ctx: *mut pvw_ctx, count: u32, zhat: *const u64, out_: *mut u64) -> c_int;
    pub fn pvw_ntt_forward_small(ctx: *mut pvw_ctx, count: u32, coeffs: *const i64, out_: *mut u64) -> c_int;
    pub fn pvw_encode_scalars(ctx: *mut pvw_ctx, count: u32, m: *const u64, out_: *mut u64) -> c_int;
    pub fn pvw_wire_layout_get(ctx: *const pvw_ctx, out_: *mut pvw_wire_layout) -> c_int;
    pub fn pvw_wire_params(ctx: *const pvw_ctx, out_: *mut u8, cap: u64) -> c_int;
    pub fn pvw_wire_ct_serialize(ctx: *mut pvw_ctx, slot0: u32, D: u32, out_: *mut u8, stride: u64, flags: u32) -> c_int;
    pub fn pvw_wire_ct_deserialize(ctx: *mut pvw_ctx, slot0: u32, D: u32, in_: *const u8, stride: u64, flags: u32) -> c_int;
    pub fn pvw_wire_pk_serialize_rows(ctx: *mut pvw_ctx, row: u32, count: u32, out_: *mut u8, flags: u32) -> c_int;
    pub fn pvw_wire_pk_deserialize_rows(ctx: *mut pvw_ctx, row: u32, count: u32, in_: *const u8, flags: u32) -> c_int;
    pub fn pvw_wire_crs_serialize(ctx: *mut pvw_ctx, out_: *mut u8, cap: u64, flags: u32) -> c_int;
    pub fn pvw_wire_crs_deserialize(ctx: *mut pvw_ctx, in_: *const u8, len: u64, flags: u32) -> c_int;
    pub fn pvw_wire_polys_serialize(ctx: *mut pvw_ctx, count: u32, polys: *const u64, out_: *mut u8) -> c_int;
    pub fn pvw_wire_polys_deserialize(ctx: *mut pvw_ctx, count: u32, in_: *const u8, polys: *mut u64) -> c_int;
    pub fn pvw_shard_export(ctx: *mut pvw_ctx, world: u32, out_: *mut pvw_shard_handle) -> c_int;
    pub fn pvw_shard_connect(ctx: *mut pvw_ctx, world: u32, rank: u32, all: *const pvw_shard_handle) -> c_int;
    pub fn pvw_shard_push_c1(ctx: *mut pvw_ctx, slot0: u32, count: u32) -> c_int;
    pub fn pvw_shard_wait_c1(ctx: *mut pvw_ctx) -> c_int;
    pub fn pvw_shard_release_c1(ctx: *mut pvw_ctx) -> c_int;
    pub fn pvw_shard_disconnect(ctx: *mut pvw_ctx) -> c_int;
    pub fn pvw_ctx_synchronize(ctx: *mut pvw_ctx) -> c_int;
    pub fn pvw_ctx_stream(ctx: *mut pvw_ctx) -> *mut c_void;
    pub fn pvw_ctx_set_option(ctx: *mut pvw_ctx, name: *const c_char, value: i64) -> c_int;
    pub fn pvw_ctx_profile(ctx: *mut pvw_ctx, kind: c_int, ms_total: *mut f64, launches: *mut u64, algorithmic_bytes: *mut f64) -> c_int;
    pub fn pvw_ctx_launch_count(ctx: *const pvw_ctx) -> u64;
    pub fn pvw_version() -> *const c_char;
}
