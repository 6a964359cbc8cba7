//! shares encrypted + decrypted per second with the UNMODIFIED pvw crate (rayon, all host cores), the structure of
//! examples/pvw.rs:130-152: `dealers` ciphertexts by encrypt (share distribution), then every party decrypts every one.
//!   cargo run --release -- <n> <k> <l> <number of 62-bit moduli> <dealers>
use std::time::Instant;

use pvw::prelude::*;
use pvw::{decrypt_party_value, encrypt};
use rand::thread_rng;
use rayon::prelude::*;

fn is_prime(n: u64) -> bool {
    if n < 2 { return false; }
    let mulmod = |a: u64, b: u64, m: u64| ((a as u128 * b as u128) % m as u128) as u64;
    let powmod = |mut a: u64, mut e: u64, m: u64| { let mut r = 1u64; a %= m; while e > 0 { if e & 1 == 1 { r = mulmod(r, a, m); } a = mulmod(a, a, m); e >>= 1; } r };
    for p in [2u64, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37] { if n % p == 0 { return n == p; } }
    let (mut d, mut s) = (n - 1, 0);
    while d % 2 == 0 { d /= 2; s += 1; }
    'outer: for a in [2u64, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37] {
        let mut x = powmod(a, d, n);
        if x == 1 || x == n - 1 { continue; }
        for _ in 1..s { x = mulmod(x, x, n); if x == n - 1 { continue 'outer; } }
        return false;
    }
    true
}

fn main() {
    let a: Vec<usize> = std::env::args().skip(1).map(|s| s.parse().unwrap()).collect();
    let (n, k, l, limbs, dealers) = (a[0], a[1], a[2], a[3], a[4]);
    // the `limbs` largest primes below 2^62 that are 1 mod 64 (bench.py uses the same set)
    let mut moduli = Vec::new();
    let mut c = ((1u64 << 62) - 1) / 64 * 64 + 1;
    while moduli.len() < limbs { if c < (1 << 62) && is_prime(c) { moduli.push(c); } c -= 64; }
    let params = PvwParametersBuilder::new().set_parties(n).set_dimension(k).set_l(l).set_moduli(&moduli).build_arc().unwrap();
    let mut rng = thread_rng();
    let crs = PvwCrs::new(&params, &mut rng).unwrap();
    let mut gpk = GlobalPublicKey::new(crs);
    let parties: Vec<Party> = (0..n).map(|i| Party::new(i, &params, &mut rng).unwrap()).collect();
    gpk.generate_all_party_keys(&parties).unwrap();
    let shares: Vec<Vec<u64>> = (0..dealers).map(|d| (0..n).map(|p| (d * 1000 + p + 1) as u64).collect()).collect();
    let t0 = Instant::now();
    let cts: Vec<_> = shares.par_iter().map(|s| encrypt(s, &gpk).unwrap()).collect();
    let t_enc = t0.elapsed();
    let t1 = Instant::now();
    let ok: usize = parties.par_iter().map(|party| {
        cts.iter().enumerate().filter(|(d, ct)| decrypt_party_value(ct, party.secret_key(), party.index()).unwrap() == shares[*d][party.index()]).count()
    }).sum();
    let t_dec = t1.elapsed();
    let total = (dealers * n) as f64;
    println!("{{\"n\": {n}, \"k\": {k}, \"l\": {l}, \"limbs\": {limbs}, \"dealers\": {dealers}, \"threads\": {}, \"encrypt_s\": {:.3}, \"decrypt_s\": {:.3}, \"shares_per_s\": {:.1}, \"recovered\": {}}}",
             rayon::current_num_threads(), t_enc.as_secs_f64(), t_dec.as_secs_f64(), total / (t_enc + t_dec).as_secs_f64(), ok as f64 / total);
}
