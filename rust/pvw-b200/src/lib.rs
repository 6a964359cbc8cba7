//! pvw-b200 -- the pvw-rs hot path on libpvw_b200.so, behind the crate's own names.
//!
//! `Device` owns one `pvw_ctx` (one GPU, one shard of parties).  Its methods have the signatures of the crate's free
//! functions (src/crypto/encryption.rs:105-296, src/crypto/decryption.rs:249-325) with `&GlobalPublicKey` replaced by
//! `&Device`, which holds A and B resident in NTT form.  Randomness is sampled here exactly where the reference samples
//! it (thread_rng(): encryption.rs:138,164,180) and handed to the library as small integers (i8 / i32).
use std::ffi::CStr;
use std::sync::{Arc, Mutex};

use fhe_math::rq::{Poly, Representation};
use num_traits::ToPrimitive;
use pvw::prelude::*;
use pvw::sampling::uniform::{sample_uniform_coefficients, sample_vec_cbd};
use pvw::PvwCiphertext;
use pvw_b200_sys as sys;

pub struct Device {
    ctx: Mutex<*mut sys::pvw_ctx>, // a pvw_ctx is single-threaded (include/pvw_b200.h); the reference types are Send + Sync
    params: Arc<PvwParameters>,
    capacity: u32,
}
unsafe impl Send for Device {}
unsafe impl Sync for Device {}

impl Drop for Device {
    fn drop(&mut self) {
        unsafe { sys::pvw_ctx_destroy(*self.ctx.lock().unwrap()) }
    }
}

fn check(ctx: *const sys::pvw_ctx, rc: i32) -> Result<()> {
    if rc == 0 {
        return Ok(());
    }
    let msg = unsafe { CStr::from_ptr(sys::pvw_last_error(ctx)) }.to_string_lossy().into_owned();
    Err(match rc {
        -1 => PvwError::InvalidParameters(msg),
        -2 => PvwError::DimensionMismatch { expected: 0, actual: 0 },
        -3 => PvwError::IndexOutOfBounds(msg),
        -4 => PvwError::EncryptionError(msg),
        -5 => PvwError::DecryptionError(msg),
        -6 => PvwError::KeyGenerationError(msg),
        -8 => PvwError::DeserializationError(msg),
        _ => PvwError::InternalError(msg),
    })
}

fn pack(polys: &[Poly]) -> Vec<u64> {
    polys.iter().flat_map(|p| p.coefficients().iter().copied()).collect() // Array2<u64> (L, l), row-major
}
fn unpack(flat: &[u64], params: &PvwParameters) -> Result<Vec<Poly>> {
    let (lc, l) = (params.moduli().len(), params.l);
    flat.chunks(lc * l)
        .map(|c| {
            let a = ndarray::Array2::from_shape_vec((lc, l), c.to_vec()).unwrap();
            Poly::try_convert_from(a, &params.context, true, Representation::Ntt).map_err(|e| PvwError::InternalError(e.to_string()))
        })
        .collect()
}
fn small_i8(v: &[i64]) -> Vec<i8> {
    v.iter().map(|&x| x as i8).collect() // CBD(variance <= 16): |x| <= 32
}

impl Device {
    /// GlobalPublicKey -> device: A and B uploaded once, `capacity` ciphertext slots reserved
    pub fn from_global_pk(gpk: &GlobalPublicKey, device: i32, capacity: u32) -> Result<Self> {
        let p = &gpk.params;
        let moduli = p.moduli().to_vec();
        let bound = |b: &num_bigint::BigInt| b.to_u64().ok_or_else(|| PvwError::InvalidParameters("error bound does not fit u64".into()));
        let desc = sys::pvw_params_desc {
            n: p.n as u32, k: p.k as u32, ell: p.l as u32, l_count: moduli.len() as u32,
            moduli: moduli.as_ptr(), psi: std::ptr::null(), // NULL: the library derives psi like fhe-math's NttOperator
            secret_variance: p.secret_variance, error_bound_1: bound(&p.error_bound_1)?, error_bound_2: bound(&p.error_bound_2)?,
            row0: 0, nrows: 0, device,
        };
        let mut ctx = std::ptr::null_mut();
        check(std::ptr::null(), unsafe { sys::pvw_ctx_create(&mut ctx, &desc) })?;
        let a: Vec<Poly> = gpk.crs.matrix.iter().cloned().collect();
        check(ctx, unsafe { sys::pvw_crs_upload(ctx, pack(&a).as_ptr(), sys::PVW_IO_HOST) })?;
        let b: Vec<Poly> = gpk.matrix.iter().cloned().collect();
        check(ctx, unsafe { sys::pvw_pk_upload_rows(ctx, 0, gpk.num_keys as u32, pack(&b).as_ptr(), sys::PVW_IO_HOST) })?;
        check(ctx, unsafe { sys::pvw_ct_reserve(ctx, capacity) })?;
        Ok(Self { ctx: Mutex::new(ctx), params: p.clone(), capacity })
    }

    /// encrypt_all_party_shares (encryption.rs:253-286): the D = n dealers in one device call; ciphertexts stay in slots 0..D
    pub fn encrypt_all_party_shares_resident(&self, all_shares: &[Vec<u64>]) -> Result<()> {
        let p = &self.params;
        let (n, k, l) = (p.n, p.k, p.l);
        if all_shares.len() != n {
            return Err(PvwError::InvalidParameters(format!("Must provide shares for all {n} parties")));
        }
        for (d, s) in all_shares.iter().enumerate() {
            if s.len() != n {
                return Err(PvwError::InvalidParameters(format!("Dealer {d} provided {} shares but needs {n}", s.len())));
            }
        }
        if n as u32 > self.capacity {
            return Err(PvwError::InvalidParameters("more dealers than reserved ciphertext slots".into()));
        }
        let mut rng = rand::thread_rng();
        let d = n;
        let mut r = Vec::with_capacity(d * k * l);
        for _ in 0..d * k {
            r.extend(sample_vec_cbd(l, p.secret_variance, &mut rng).map_err(|e| PvwError::SamplingError(e.to_string()))?);
        }
        let small = |bound: &num_bigint::BigInt, count: usize, rng: &mut rand::rngs::ThreadRng| -> Result<Vec<i32>> {
            Ok(sample_uniform_coefficients(bound, count, rng).iter().map(|x| x.to_i32().unwrap_or(0)).collect())
        };
        let e1 = small(&p.error_bound_1, d * k * l, &mut rng)?;
        let e2 = small(&p.error_bound_2, d * n * l, &mut rng)?;
        let m: Vec<u64> = all_shares.iter().flatten().copied().collect();
        let r8 = small_i8(&r);
        let ctx = *self.ctx.lock().unwrap();
        check(ctx, unsafe {
            sys::pvw_encrypt_batch(ctx, 0, d as u32, 0, d as u32, m.as_ptr(), r8.as_ptr().cast(), e1.as_ptr().cast(), e2.as_ptr().cast(),
                                   sys::PVW_IO_HOST | sys::PVW_IN_SECRET_I8 | sys::PVW_IN_ERROR_I32)
        })
    }

    /// ... and as owned `PvwCiphertext` values, like the reference returns them
    pub fn encrypt_all_party_shares(&self, all_shares: &[Vec<u64>]) -> Result<Vec<PvwCiphertext>> {
        self.encrypt_all_party_shares_resident(all_shares)?;
        (0..all_shares.len() as u32).map(|slot| self.ciphertext(slot)).collect()
    }

    pub fn ciphertext(&self, slot: u32) -> Result<PvwCiphertext> {
        let p = &self.params;
        let poly = p.moduli().len() * p.l;
        let (mut c1, mut c2) = (vec![0u64; p.k * poly], vec![0u64; p.n * poly]);
        let ctx = *self.ctx.lock().unwrap();
        check(ctx, unsafe { sys::pvw_ct_download(ctx, slot, c1.as_mut_ptr(), c2.as_mut_ptr()) })?;
        Ok(PvwCiphertext { c1: unpack(&c1, p)?, c2: unpack(&c2, p)?, params: p.clone() })
    }

    /// decrypt_party_shares (decryption.rs:281-325) on the ciphertexts resident in slots 0..n
    pub fn decrypt_party_shares(&self, secret_key: &SecretKey, party_index: usize) -> Result<Vec<u64>> {
        let n = self.params.n;
        if party_index >= n {
            return Err(PvwError::InvalidParameters(format!("Party index {party_index} exceeds maximum {}", n - 1)));
        }
        let sk: Vec<i8> = secret_key.coefficients().iter().flatten().map(|&x| x as i8).collect();
        let mut out = vec![0u64; n];
        let idx = [party_index as u32];
        let ctx = *self.ctx.lock().unwrap();
        check(ctx, unsafe {
            sys::pvw_decrypt_batch(ctx, n as u32, std::ptr::null(), 1, idx.as_ptr(), sk.as_ptr().cast(), out.as_mut_ptr(),
                                   sys::PVW_IO_HOST | sys::PVW_IN_SECRET_I8)
        })?;
        Ok(out)
    }

    /// decrypt_party_value (decryption.rs:249-278) on a ciphertext that lives on the host: uploaded to a slot first
    pub fn decrypt_party_value(&self, ciphertext: &PvwCiphertext, secret_key: &SecretKey, party_index: usize) -> Result<u64> {
        let ctx = *self.ctx.lock().unwrap();
        let slot = self.capacity - 1;
        check(ctx, unsafe { sys::pvw_ct_upload(ctx, slot, pack(&ciphertext.c1).as_ptr(), pack(&ciphertext.c2).as_ptr()) })?;
        let sk: Vec<i8> = secret_key.coefficients().iter().flatten().map(|&x| x as i8).collect();
        let (idx, slots, mut out) = ([party_index as u32], [slot], [0u64]);
        check(ctx, unsafe {
            sys::pvw_decrypt_batch(ctx, 1, slots.as_ptr(), 1, idx.as_ptr(), sk.as_ptr().cast(), out.as_mut_ptr(),
                                   sys::PVW_IO_HOST | sys::PVW_IN_SECRET_I8)
        })?;
        Ok(out[0])
    }

    /// generate_all_party_keys (src/keys/public_key.rs:376-401) on the device: rows land in B at party.index()
    pub fn generate_all_party_keys(&self, parties: &[Party]) -> Result<()> {
        let p = &self.params;
        let mut rng = rand::thread_rng();
        let ctx = *self.ctx.lock().unwrap();
        for party in parties {
            let sk: Vec<i8> = party.secret_key().coefficients().iter().flatten().map(|&x| x as i8).collect();
            let e: Vec<i32> = sample_uniform_coefficients(&p.error_bound_1, p.k * p.l, &mut rng).iter().map(|x| x.to_i32().unwrap_or(0)).collect();
            check(ctx, unsafe {
                sys::pvw_keygen_batch(ctx, party.index() as u32, 1, sk.as_ptr().cast(), e.as_ptr().cast(),
                                      sys::PVW_IO_HOST | sys::PVW_IN_SECRET_I8 | sys::PVW_IN_ERROR_I32)
            })?;
        }
        Ok(())
    }
}
