//! NOT part of the crate build (no `mod` line points here): the bodies a maintainer pastes into
//! fawkes-crypto/src/backend/bellman_groth16/ behind `#[cfg(feature = "backend_b200_groth16")]`.  Every public
//! type and signature of the module stays as it is; only what sits behind the three bellman calls changes.
//!
//! ---------------------------------------------------------------------------------------------- mod.rs
//! `Parameters` keeps its four fields (mod.rs:139) but `.0` holds the bellman BYTES instead of the parsed
//! `bellman::groth16::Parameters`; read/write keep their framing (mod.rs:150-175).
//!
//! pub struct Parameters<E: Engine>(pub Vec<u8>, pub u32, pub Vec<u8>, pub BitVec,
//!                                  pub(crate) b200::KeyCell, pub(crate) (bool, bool), PhantomData<E>);
//!
//! pub fn read(reader: &mut &[u8], disallow_points_at_infinity: bool, checked: bool) -> std::io::Result<Self> {
//!     let e1 = BorshDeserialize::deserialize(reader)?;                 // num_gates      (mod.rs:160)
//!     let e2 = BorshDeserialize::deserialize(reader)?;                 // gate blob      (mod.rs:161)
//!     let e3_len = <u32 as BorshDeserialize>::deserialize(reader)? as usize;
//!     let e3_buf: Vec<u8> = BorshDeserialize::deserialize(reader)?;
//!     if e3_len > e3_buf.len() * 8 { return Err(io::Error::new(InvalidData, "inconsistent bitvec length")); }
//!     let mut e3 = BitVec::from_bytes(&e3_buf); e3.truncate(e3_len);
//!     let e0 = reader.to_vec();                                        // bellman bytes, validated at key load
//!     *reader = &reader[reader.len()..];
//!     let p = Self(e0, e1, e2, e3, Default::default(), (disallow_points_at_infinity, checked), PhantomData);
//!     p.b200_key()?;   // eager load keeps read()'s contract: an invalid key is an io::Error HERE (mod.rs:173)
//!     Ok(p)
//! }
//! pub fn get_vk(&self) -> verifier::VK<E> { verifier::VK::from_bellman_bytes(&self.0) }   // first 580 + 64 n_ic bytes
//!
//! fn b200_key(&self) -> io::Result<&fawkes_b200::ProvingKey<'static>> {   // cached in the KeyCell
//!     self.4.get_or_try_init(|| fawkes_b200::ProvingKey::load(b200::ctx(), &fawkes_b200::Parameters {
//!         bellman_bytes: self.0.clone(), num_gates: self.1, gates_brotli: self.2.clone(),
//!         disallow_points_at_infinity: (self.5).0, checked: (self.5).1 }))
//! }
//!
//! ---------------------------------------------------------------------------------------------- prover.rs
//! pub fn prove_with_rs<'a, E, Pub, Sec, C>(params: &'a Parameters<E>, input_pub: &Pub::Value, input_sec: &Sec::Value,
//!         circuit: C, r: Num<E::Fr>, s: Num<E::Fr>) -> (Vec<Num<E::Fr>>, Proof<E>)
//! where E: Engine, Pub: Signal<WitnessCS<'a, E::Fr>>, Sec: Signal<WitnessCS<'a, E::Fr>>, C: Fn(Pub, Sec) {
//!     let ref rcs = params.get_witness_rcs();                          // prover.rs:69
//!     let signal_pub = Pub::alloc(rcs, Some(input_pub));
//!     signal_pub.inputize();
//!     let signal_sec = Sec::alloc(rcs, Some(input_sec));
//!     circuit(signal_pub, signal_sec);                                 // prover.rs:74: witness generation, host
//!     let cs = rcs.borrow();
//!     assert!(cs.const_tracker_index == cs.const_tracker.len(), "not all cached data used");   // prover.rs:83
//!     // Num<Fr> is #[repr(transparent)] over 4 x u64 Montgomery limbs (ff-uint/src/num/mod.rs:21-23)
//!     let vi: &[[u64; 4]] = unsafe { std::slice::from_raw_parts(cs.values_input.as_ptr() as *const _, cs.values_input.len()) };
//!     let va: &[[u64; 4]] = unsafe { std::slice::from_raw_parts(cs.values_aux.as_ptr() as *const _, cs.values_aux.len()) };
//!     let raw = params.b200_key().unwrap().prove_with_rs(vi, va, unsafe { &*(&r as *const _ as *const [u64; 4]) },
//!                                                        unsafe { &*(&s as *const _ as *const [u64; 4]) });
//!     let proof: Proof<E> = unsafe { std::ptr::read(raw.as_ptr() as *const Proof<E>) };   // same bytes (prover.rs:13-17)
//!     let inputs = cs.values_input[1..].to_vec();                      // prover.rs:84-87
//!     (inputs, proof)
//! }
//! pub fn prove<..>(params, input_pub, input_sec, circuit) -> (Vec<Num<E::Fr>>, Proof<E>) {   // prover.rs:63-90
//!     let mut rng = fawkes_b200::osrng::OsRng::new();                  // the reference's sampling convention
//!     let (r, s) = (Num::from_mont_uint_unchecked(rng.gen_fr()), Num::from_mont_uint_unchecked(rng.gen_fr()));
//!     prove_with_rs(params, input_pub, input_sec, circuit, r, s)
//! }
//!
//! ---------------------------------------------------------------------------------------------- setup.rs
//! After `circuit(signal_pub, signal_sec);` and the gate serialisation of setup.rs:25-32 (unchanged):
//!     let bytes = fawkes_b200::setup(b200::ctx(), &gates_brotli, num_gates, cs.num_input() as u32, cs.num_aux() as u32);
//!     Parameters(bytes, num_gates, gates_brotli, cs.const_tracker.clone(), Default::default(), (false, false), PhantomData)
//!
//! ---------------------------------------------------------------------------------------------- verifier.rs
//! pub fn verify<E: Engine>(vk: &VK<E>, proof: &Proof<E>, inputs: &[Num<E::Fr>]) -> bool {     // verifier.rs:75-81
//!     let mut vk_raw = Vec::with_capacity(448 + 64 * vk.ic.len());
//!     vk_raw.extend_from_slice(as_bytes(&vk.alpha)); vk_raw.extend_from_slice(as_bytes(&vk.beta));
//!     vk_raw.extend_from_slice(as_bytes(&vk.gamma)); vk_raw.extend_from_slice(as_bytes(&vk.delta));
//!     for p in &vk.ic { vk_raw.extend_from_slice(as_bytes(p)); }
//!     fawkes_b200::verify(&vk_raw, unsafe { &*(proof as *const _ as *const [u8; 256]) },
//!                         unsafe { std::slice::from_raw_parts(inputs.as_ptr() as *const [u64; 4], inputs.len()) })
//! }
