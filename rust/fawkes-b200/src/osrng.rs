//! Blinding-factor sampling with the reference's conventions, so that `prove()` draws r and s exactly the way
//! `bellman::groth16::create_random_proof(bcs, &params.0, rng)` does with the reference's `OsRng`
//! (fawkes-crypto/src/backend/bellman_groth16/osrng.rs:12-18, prover.rs:78-80):
//!   * `OsRng::next_u32` = 4 bytes from `getrandom`, read BIG-endian (osrng.rs:13-17); it is the only method the
//!     reference implements, so `next_u64` is rand 0.4's default: `(next_u32 as u64) << 32 | next_u32 as u64`
//!     (high word first);
//!   * ff_ce's `impl Rand for Fr` [restated, crate not vendored]: fill the 4 limbs with `next_u64` in limb order,
//!     clear the top `REPR_SHAVE_BITS` (= 2 for BN254 Fr) bits of the top limb, accept when the value is below the
//!     modulus; the limbs are taken AS THEY ARE as the Montgomery representation (no conversion).
//! The result is therefore already the 4 x u64 Montgomery `Num<Fr>` that `fb_prove` takes for r and s.
use getrandom::getrandom;

/// BN254 scalar field modulus r, little-endian limbs (fawkes-crypto/src/engines/bn256/mod.rs:13).
pub const FR_MODULUS: [u64; 4] = [0x43e1f593f0000001, 0x2833e84879b97091, 0xb85045b68181585d, 0x30644e72e131a029];
const REPR_SHAVE_BITS: u32 = 2;

pub struct OsRng;

impl OsRng {
    pub fn new() -> Self {
        OsRng
    }
    pub fn next_u32(&mut self) -> u32 {
        let mut buf = [0u8; 4];
        getrandom(&mut buf).expect("getrandom");
        u32::from_be_bytes(buf)
    }
    pub fn next_u64(&mut self) -> u64 {
        ((self.next_u32() as u64) << 32) | (self.next_u32() as u64)
    }
    /// One uniformly random `Num<Fr>` as raw Montgomery limbs.
    pub fn gen_fr(&mut self) -> [u64; 4] {
        loop {
            let mut l = [0u64; 4];
            for x in l.iter_mut() {
                *x = self.next_u64();
            }
            l[3] &= u64::MAX >> REPR_SHAVE_BITS;
            if lt(&l, &FR_MODULUS) {
                return l;
            }
        }
    }
}

fn lt(a: &[u64; 4], b: &[u64; 4]) -> bool {
    for i in (0..4).rev() {
        if a[i] != b[i] {
            return a[i] < b[i];
        }
    }
    false
}
