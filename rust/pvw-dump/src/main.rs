//! Prints known answers of the hot path from the UNMODIFIED pvw crate, one JSON file per parameter set.
//!
//! Everything below goes through the crate's public API and fhe-math's `Poly` operators -- the same calls
//! `encrypt` / `decrypt_party_value` make (src/crypto/encryption.rs:147-200, src/crypto/decryption.rs:249-278) --
//! with the three `thread_rng()` draws of `encrypt` replaced by fixed small integers, so that the output is
//! reproducible.  tests/test_from_crate.py compares the oracle and the CUDA path with these files.
//!
//!   cargo run --release -- <out_dir>
use std::{fs, path::Path, sync::Arc};

use fhe_math::rq::{Poly, Representation};
use fhe_traits::Serialize as FheSerialize;
use num_bigint::{BigInt, BigUint};
use pvw::prelude::*;
use pvw::{PvwCiphertext, decrypt_party_value};
use serde_json::{Value, json};

/// deterministic small integers in [-bound, bound]: splitmix64 of (tag, index), the same stream as oracle/pvw_oracle.py
fn splitmix64(mut x: u64) -> u64 {
    x = x.wrapping_add(0x9E3779B97F4A7C15);
    let mut z = x;
    z = (z ^ (z >> 30)).wrapping_mul(0xBF58476D1CE4E5B9);
    z = (z ^ (z >> 27)).wrapping_mul(0x94D049BB133111EB);
    z ^ (z >> 31)
}
fn small(tag: u64, index: u64, bound: i64) -> i64 {
    let u = splitmix64(0x5056572D42323030 ^ (tag << 56) ^ index);
    (((u as u128) * ((2 * bound + 1) as u128)) >> 64) as i64 - bound
}

fn residues(p: &Poly) -> Value {
    // Poly.coefficients: Array2<u64> (L, l), row-major (src/params/parameters.rs:455-458)
    let rows: Vec<Vec<String>> = p.coefficients().outer_iter().map(|r| r.iter().map(|x| x.to_string()).collect()).collect();
    json!(rows)
}
fn from_small(coeffs: &[i64], params: &Arc<PvwParameters>) -> Poly {
    let mut p = Poly::from_coefficients(coeffs, &params.context).expect("from_coefficients");
    p.change_representation(Representation::Ntt);
    p
}
fn error_poly(coeffs: &[i64], params: &Arc<PvwParameters>) -> Poly {
    // sample_error_1/2 (parameters.rs:264-284) with the sampled BigInts replaced by `coeffs`
    let big: Vec<BigInt> = coeffs.iter().map(|&c| BigInt::from(c)).collect();
    let mut p = params.bigints_to_poly(&big).expect("bigints_to_poly");
    p.change_representation(Representation::Ntt);
    p
}

fn dump(name: &str, n: usize, k: usize, l: usize, moduli: &[u64], variance: f32, b1: u32, b2: u32, out: &Path) {
    let params = PvwParametersBuilder::new()
        .set_parties(n).set_dimension(k).set_l(l).set_moduli(moduli)
        .set_secret_variance(variance).set_error_bounds_u32(b1, b2)
        .build_arc().expect("parameters");

    // psi and slot order: NTT(X) per modulus; slot 0 is psi itself (slot i = psi^(2 brv(i) + 1))
    let mut x = vec![0i64; l];
    x[1] = 1;
    let ntt_x = from_small(&x, &params);
    let mut ramp: Vec<i64> = (0..l as i64).map(|i| i - 3).collect();
    ramp[0] = -7;
    let ntt_ramp = from_small(&ramp, &params);

    // seeded CRS (src/params/crs.rs:45-67)
    let seed = [0x42u8; 32];
    let crs = PvwCrs::new_deterministic(&params, seed).expect("crs");
    let crs_tag = PvwCrs::new_from_tag(&params, "pvw-b200 parity").expect("crs from tag");

    // keys with fixed coefficients; public rows WITHOUT error through the crate's own multiply_by_secret_key
    // (src/params/crs.rs:138-171), plus a fixed error added with Poly + (public_key.rs:124-139)
    let mut sks = Vec::new();
    let mut b_rows: Vec<Vec<Poly>> = Vec::new();
    let mut ke_all: Vec<Vec<Vec<i64>>> = Vec::new();
    for p in 0..n {
        let coeffs: Vec<Vec<i64>> = (0..k).map(|j| (0..l).map(|t| small(1, ((p * k + j) * l + t) as u64, 1)).collect()).collect();
        let sk = SecretKey::from_coefficients(params.clone(), coeffs).expect("sk");
        let sa = crs.multiply_by_secret_key(&sk).expect("sA");
        let ke: Vec<Vec<i64>> = (0..k).map(|c| (0..l).map(|t| small(2, ((p * k + c) * l + t) as u64, b1 as i64)).collect()).collect();
        let row: Vec<Poly> = sa.iter().zip(ke.iter()).map(|(a, e)| a + &error_poly(e, &params)).collect();
        sks.push(sk);
        b_rows.push(row);
        ke_all.push(ke);
    }
    let mut gpk = GlobalPublicKey::new(crs.clone());
    for (p, row) in b_rows.iter().enumerate() {
        let pk = PublicKey { key_polynomials: row.clone(), params: params.clone() };
        gpk.add_public_key(p, pk).expect("add_public_key");
    }

    // one ciphertext with fixed r, e1, e2: the body of encrypt (encryption.rs:147-200)
    let r: Vec<Vec<i64>> = (0..k).map(|j| (0..l).map(|t| small(3, (j * l + t) as u64, 1)).collect()).collect();
    let e1: Vec<Vec<i64>> = (0..k).map(|j| (0..l).map(|t| small(4, (j * l + t) as u64, b1 as i64)).collect()).collect();
    let e2: Vec<Vec<i64>> = (0..n).map(|p| (0..l).map(|t| small(5, (p * l + t) as u64, b2 as i64)).collect()).collect();
    let m: Vec<u64> = (0..n as u64).map(|p| 1000 + p + 1).chain(std::iter::empty()).collect();
    let r_polys: Vec<Poly> = r.iter().map(|c| from_small(c, &params)).collect();
    let mut c1 = crs.multiply_by_randomness(&r_polys).expect("A r");
    for (c, e) in c1.iter_mut().zip(e1.iter()) {
        *c = &*c + &error_poly(e, &params);
    }
    let mut c2 = Vec::new();
    for p in 0..n {
        let mut acc = Poly::zero(&params.context, Representation::Ntt);
        for (j, rj) in r_polys.iter().enumerate() {
            let prod = gpk.get_polynomial(p, j).unwrap() * rj;
            acc = &acc + &prod;
        }
        let enc = params.encode_scalar(m[p] as i64).expect("encode");
        c2.push(&(&acc + &enc) + &error_poly(&e2[p], &params));
    }
    let ct = PvwCiphertext { c1: c1.clone(), c2: c2.clone(), params: params.clone() };
    let dec: Vec<u64> = (0..n).map(|p| decrypt_party_value(&ct, &sks[p], p).expect("decrypt")).collect();

    let delta: &BigUint = params.delta();
    let doc = json!({
        "set": name, "n": n, "k": k, "l": l, "moduli": moduli.iter().map(|q| q.to_string()).collect::<Vec<_>>(),
        "secret_variance": variance, "error_bound_1": b1, "error_bound_2": b2,
        "delta": delta.to_string(), "delta_power_l_minus_1": params.delta_power_l_minus_1().to_string(),
        "correctness_condition": params.verify_correctness_condition(),
        "ntt_x": residues(&ntt_x), "ramp": ramp, "ntt_ramp": residues(&ntt_ramp),
        "encode_scalar_12345": residues(&params.encode_scalar(12345).unwrap()),
        "encode_scalar_minus_7": residues(&params.encode_scalar(-7).unwrap()),
        "crs_seed_hex": seed.iter().map(|b| format!("{b:02x}")).collect::<String>(),
        "crs_a_0_0": residues(crs.get(0, 0).unwrap()), "crs_a_last": residues(crs.get(k - 1, k - 1).unwrap()),
        "crs_tag": "pvw-b200 parity", "crs_tag_a_0_0": residues(crs_tag.get(0, 0).unwrap()),
        "sk": sks.iter().map(|s| s.coefficients().to_vec()).collect::<Vec<_>>(), "key_error": ke_all,
        "b_row_0": b_rows[0].iter().map(residues).collect::<Vec<_>>(),
        "b_row_last": b_rows[n - 1].iter().map(residues).collect::<Vec<_>>(),
        "r": r, "e1": e1, "e2": e2, "m": m.iter().map(|x| x.to_string()).collect::<Vec<_>>(),
        "c1": c1.iter().map(residues).collect::<Vec<_>>(), "c2": c2.iter().map(residues).collect::<Vec<_>>(),
        "decrypted": dec.iter().map(|x| x.to_string()).collect::<Vec<_>>(),
        "poly_to_bytes_hex_c1_0": c1[0].to_bytes().iter().map(|b| format!("{b:02x}")).collect::<String>(),
        "ciphertext_bincode_hex": bincode::serialize(&ct).unwrap().iter().map(|b| format!("{b:02x}")).collect::<String>(),
        "params_bincode_hex": bincode::serialize(&*params).unwrap().iter().map(|b| format!("{b:02x}")).collect::<String>(),
    });
    fs::create_dir_all(out).unwrap();
    fs::write(out.join(format!("{name}.json")), serde_json::to_string(&doc).unwrap()).unwrap();
    println!("{name}: wrote {} (decrypted == m: {})", out.join(format!("{name}.json")).display(), dec == m);
}

fn main() {
    let out = std::env::args().nth(1).unwrap_or_else(|| "../../tests/golden/from_crate".into());
    let out = Path::new(&out);
    // examples/pvw.rs:28-32 and tests/crypto.rs:236-305
    dump("EX", 7, 32, 8, &[0xffffc4001, 0x1ffffe0001], 0.5, 50, 50, out);
    dump("T16", 10, 4, 16, &[0xffffee001, 0xffffc4001, 0x1ffffe0001], 0.5, 50, 50, out);
    // examples/pvw_valid_dec.rs:40-52 with k cut to 40 (the file stays small)
    dump("VDs", 6, 40, 8, &[0x800000022a0001, 0x800000021a0001, 0x80000002120001, 0x80000001f60001], 10.0, 1, 1172385, out);
}
